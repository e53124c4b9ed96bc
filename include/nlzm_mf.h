/* nlzm_mf.h — C ABI of the B200-native match-finding engine for NLZM.
 *
 * Drop-in boundary. The reference (nauful/NLZM 1.03, NLZM.cpp) has no plugin API:
 * its encoder owns four concrete finder objects and calls them by name. Each entry point below
 * names the reference interface it replaces:
 *
 *   nlzm_mf_create      <- MatchFinderHT::Init x2, MatchFinderBT::Init, MatchFinderRK256::Init
 *                          as called from encode_file (NLZM.cpp:1745-1753), incl. the hist_bits
 *                          shrink rule (NLZM.cpp:1716-1718) and the -window clamp (NLZM.cpp:2085)
 *   nlzm_mf_set_input*  <- the RingDictionary the finders read (NLZM.cpp:754-764, 1729-1738):
 *                          hist ring + lookahead are replaced by the flat input resident in HBM
 *   nlzm_mf_find        <- the per-position calls ht2/ht3/bt4/rk.FindAndUpdate(mt, h, P, dict)
 *                          in parse_table (NLZM.cpp:1514-1541) for a whole range of positions,
 *                          and the X.Shift(W) calls (NLZM.cpp:1786-1792), which become part of
 *                          the closed-form geometry
 *   nlzm_mf_view steps  <- the MatchTable::Update(delta, len) calls the finders make
 *                          (NLZM.cpp:835-852): one step == one Update, same pre-conditions
 *   nlzm_mf_destroy     <- X.Release() (NLZM.cpp:1901-1904)
 *
 * Error convention: every call returns 0 on success or a negative nlzm_mf_status / positive CUDA
 * error code; nlzm_mf_last_error() gives the text. The library never aborts the process (the
 * reference's ASSERT -> exit(-1), NLZM.cpp:25, is the host shim's business).
 *
 * Semantics ("R2"): the steps returned for position a are exactly what the
 * reference finders report when BT4's test cap is lifted and every finder is called at every
 * eligible position — a pure function of (bytes, hist_bits).
 */
#ifndef NLZM_MF_H
#define NLZM_MF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLZM_MF_ABI_VERSION 2

enum nlzm_mf_finder {
    NLZM_MF_HT2 = 1,      /* MatchFinderHT, 2-byte hash, 1 row   (NLZM.cpp:1750) */
    NLZM_MF_HT3 = 2,      /* MatchFinderHT, 3-byte hash, 2 rows  (NLZM.cpp:1751) */
    NLZM_MF_BT4 = 4,      /* MatchFinderBT, exhaustive           (NLZM.cpp:1752) */
    NLZM_MF_RK256 = 8,    /* MatchFinderRK256                    (NLZM.cpp:1753) */
    NLZM_MF_ALL = 15
};

enum nlzm_mf_status {
    NLZM_MF_OK = 0,
    NLZM_MF_E_ARG = -1,        /* bad argument */
    NLZM_MF_E_NOMEM = -2,      /* host or device allocation failed */
    NLZM_MF_E_STATE = -3,      /* call order violated (e.g. find before set_input) */
    NLZM_MF_E_NODEVICE = -4,   /* no usable CUDA device: there is no CPU fallback */
    NLZM_MF_E_OVERFLOW = -5    /* internal candidate buffer overflow (retry is automatic; see last_error) */
};

typedef struct nlzm_mf nlzm_mf;

typedef struct {
    uint32_t struct_size;      /* sizeof(nlzm_mf_config) */
    uint32_t hist_bits;        /* as given to -window:N (clamped 15..28, then shrunk to the file) */
    uint64_t file_len;         /* total input length in bytes (< 2^31) */
    int32_t device;            /* CUDA device ordinal */
    uint32_t finder_mask;      /* nlzm_mf_finder bits; 0 = all */
    uint64_t max_range;        /* largest end-begin a find call will use (<= 2^28); 0 = file_len */
} nlzm_mf_config;

/* The geometry the reference derives from (file_len, hist_bits); part of the matcher semantics. */
typedef struct {
    uint32_t hist_bits, window, frame_bits, chunk_size, feed_size;
    uint32_t ht2_bits, ht3_bits, bt4_bits, rk_bits;
} nlzm_mf_geometry;

/* One staircase step == one MatchTable::Update(dist, len). Six bytes (three u16, no padding): the
 * candidate records are what crosses PCIe, 3.5 of them per input byte on text. */
typedef struct {
    uint16_t dist_lo, dist_hi;     /* distance = dist_lo | dist_hi << 16 */
    uint16_t len;
} nlzm_mf_step;
#define NLZM_MF_STEP_DIST(s) ((uint32_t)(s).dist_lo | ((uint32_t)(s).dist_hi << 16))

/* Candidates of positions [begin, end): position a owns steps[offsets[a-begin] .. offsets[a-begin+1]),
 * strictly increasing in len and dist. Pointers are HOST pointers into pinned memory owned by the
 * engine (or DEVICE pointers for nlzm_mf_find_device), valid until the next find on the same slot. */
typedef struct {
    uint64_t begin, end;
    uint64_t n_steps;
    const uint32_t *offsets;       /* end - begin + 1 entries */
    const nlzm_mf_step *steps;
} nlzm_mf_view;

typedef struct {
    uint64_t kernel_launches;      /* kernels launched by this engine since creation */
    uint64_t tuples_last;          /* candidate tuples produced by the last find */
    float ms_rank, ms_levels, ms_ht, ms_rk, ms_merge, ms_total, ms_d2h;  /* last find, CUDA events */
} nlzm_mf_stats;

int nlzm_mf_abi_version(void);
int nlzm_mf_get_geometry(uint64_t file_len, uint32_t hist_bits, nlzm_mf_geometry *out);

int nlzm_mf_create(const nlzm_mf_config *cfg, nlzm_mf **out);
void nlzm_mf_destroy(nlzm_mf *mf);
const char *nlzm_mf_last_error(const nlzm_mf *mf);   /* mf may be NULL: error of the last failed create */

/* Input: the whole file is kept resident in HBM (it is replicated on every GPU of a box). */
int nlzm_mf_set_input(nlzm_mf *mf, const uint8_t *host_data, uint64_t len);          /* H2D copy */
int nlzm_mf_set_input_device(nlzm_mf *mf, const void *device_data, uint64_t len);    /* D2D copy */

/* Two result slots (0, 1) allow find(N+1) to overlap the host's consumption of slot N. */
int nlzm_mf_find(nlzm_mf *mf, uint64_t begin, uint64_t end, int slot, nlzm_mf_view *out);        /* sync, host view */
int nlzm_mf_find_device(nlzm_mf *mf, uint64_t begin, uint64_t end, int slot, nlzm_mf_view *out); /* sync, device view */
int nlzm_mf_submit(nlzm_mf *mf, uint64_t begin, uint64_t end, int slot);                         /* async enqueue */
int nlzm_mf_fetch(nlzm_mf *mf, int slot, nlzm_mf_view *out);                                     /* wait + host view */

int nlzm_mf_get_stats(const nlzm_mf *mf, nlzm_mf_stats *out);

/* Tuning / test knobs (no reference counterpart; results never depend on them):
 *   "ht_margin"      positions before a range for which the HT stage materialises per-position data
 *                    (default: the whole prefix, 12 bytes per position; smaller = less memory, slower look-ups)
 *   "ht_coarse_log"  log2 of the coarse table spacing of the far prefix (default 20) */
int nlzm_mf_set_option(nlzm_mf *mf, const char *key, uint64_t value);

/* Measurement aid (no reference counterpart): with profiling enabled every kernel launch is
 * bracketed by CUDA events on its stream and accumulated per kernel name (process wide). */
typedef struct {
    char name[56];
    uint64_t launches;
    double ms;
} nlzm_mf_kernel_time;
int nlzm_mf_profile(int enable);                                    /* enable=0 also clears the table */
int nlzm_mf_get_kernel_times(nlzm_mf_kernel_time *out, uint32_t cap, uint32_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* NLZM_MF_H */
