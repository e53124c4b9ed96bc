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

#define NLZM_MF_ABI_VERSION 3

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
 * candidate records are what crosses PCIe, 3.5 of them per input byte on text. The spare bits carry what
 * the parser needs before it can price the candidate (NLZM.cpp:1556-1596), computed on the GPU:
 *   dist_lo            distance bits 0..15
 *   dist_hi  0..11     distance bits 16..27        12..13  shortest length allowed at this distance - 2
 *                                                          (get_match_min, NLZM.cpp:813-821)
 *   len      0..8      length (2..264)             9..14   distance slot of the stream's distance code
 *                                                          (NLZM.cpp:1219-1236; raw bits = slot < 4 ? 0 : slot/2 - 1) */
typedef struct {
    uint16_t dist_lo, dist_hi;
    uint16_t len;
} nlzm_mf_step;
#define NLZM_MF_STEP_DIST(s) ((uint32_t)(s).dist_lo | (((uint32_t)(s).dist_hi & 0x0FFFu) << 16))
#define NLZM_MF_STEP_LEN(s) ((uint32_t)(s).len & 0x1FFu)
#define NLZM_MF_STEP_SLOT(s) (((uint32_t)(s).len >> 9) & 0x3Fu)
#define NLZM_MF_STEP_SHORTEST(s) (2u + (((uint32_t)(s).dist_hi >> 12) & 3u))

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
    float ms_cross;                /* last find: queries against retained / imported segments */
    float ms_prepare;              /* last nlzm_mf_prepare (also counted in the ms_total of the find that continues it) */
    uint32_t segments_queried;     /* last find: retained / imported segments its first block looked into */
    uint32_t segments_retained;    /* segments kept after the last find */
    float ms_import;               /* device time of the segment copies since the last nlzm_mf_prepare (CUDA events) */
    uint32_t reserved;
    uint64_t bytes_imported;       /* ... and their size */
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

/* Matcher state that outlives a call. The reference's BT4 tree persists from one chunk to the next
 * (MatchFinderBT heads/tree, NLZM.cpp:959-972, shifted but never rebuilt, NLZM.cpp:1024-1031). Here the
 * persistent state is a list of SEGMENTS: the sorted level array and greater-position pointers of the
 * last window's worth of positions, kept in HBM after every find. A find whose range starts where
 * retained segments end queries them instead of re-ranking and re-merging the window behind its
 * range; any other range is computed from scratch (results are identical either way).
 *
 * Position sharding across GPUs (the input is replicated, each engine owns a range):
 *   1. every engine:  nlzm_mf_prepare(own_begin, own_end)      rank + merge its own range only
 *   2. every engine:  nlzm_mf_export_segments(...)              descriptors (device pointers; same process), or
 *                     nlzm_mf_publish_segments(...)             copies in the engine's export buffer + its CUDA IPC handle
 *   3. engine r:      nlzm_mf_import_segment(...) for the segments of the ranges behind own_begin that lie
 *                     within the window (copied over NVLink: peer copy in one process, CUDA IPC across processes)
 *   4. engine r:      nlzm_mf_find(own_begin, own_end, ...)     continues from step 1
 * No collective is involved; step 3 is a one-sided read of the neighbours' HBM. */
typedef struct {
    uint64_t pos_begin, pos_end;            /* absolute positions covered */
    uint64_t origin;                        /* element positions are relative to this absolute offset */
    uint64_t n_elems;
    uint64_t elems_offset_bytes, elems_bytes;   /* the segment's slice inside the exporting engine's allocations */
    uint64_t ptrs_offset_bytes, ptrs_bytes;
    const void *elems_alloc, *ptrs_alloc;   /* device pointers (same-process import) */
    int32_t device;
    uint32_t flags;                         /* bit 0 / 1: ipc_elems / ipc_ptrs valid */
    uint8_t ipc_elems[64], ipc_ptrs[64];    /* cudaIpcMemHandle_t of the two allocations (import from another process) */
} nlzm_mf_segment;
int nlzm_mf_prepare(nlzm_mf *mf, uint64_t begin, uint64_t end);
int nlzm_mf_export_segments(nlzm_mf *mf, nlzm_mf_segment *out, uint32_t cap, uint32_t *n_out);
/* via: 0 = elems_alloc / ptrs_alloc are device pointers of this process (peer copy), 1 = open the IPC handles,
 *      2 = elems_alloc / ptrs_alloc are HOST copies of the two slices made with nlzm_mf_read_segment;
 *      | 0x100 (with 0 or 1): asynchronous — the copies are queued on the engine's copy stream and the call
 *      returns; the next find runs its HT and RK stages while they arrive and waits for them before it
 *      queries the segments. The exporter must keep the source unchanged until that find has returned. */
/* Export for OTHER PROCESSES: copies this engine's own segments, cut down to positions >= from_pos, into one export
 * buffer that lives as long as the engine and describes the copies (call with out == NULL to get the count). The
 * descriptors carry the CUDA IPC handle of that buffer: an importer maps it once, however often it is refilled. */
int nlzm_mf_publish_segments(nlzm_mf *mf, uint64_t from_pos, nlzm_mf_segment *out, uint32_t cap, uint32_t *n_out);
int nlzm_mf_import_segment(nlzm_mf *mf, const nlzm_mf_segment *seg, int via);
int nlzm_mf_read_segment(nlzm_mf *mf, uint32_t index, void *elems_host, void *ptrs_host);
int nlzm_mf_drop_segments(nlzm_mf *mf);
/* Keep only positions >= from_pos in the retained list (a straddling segment is cut down to them, order kept): what a
 * neighbour then imports is at most one window. */
int nlzm_mf_trim_segments(nlzm_mf *mf, uint64_t from_pos);

/* Tuning / test knobs (no reference counterpart; results never depend on them):
 *   "ht_margin"      positions before a range for which the HT stage materialises per-position data
 *                    (default: the whole prefix, 12 bytes per position; smaller = less memory, slower look-ups)
 *   "ht_coarse_log"  log2 of the coarse table spacing of the far prefix (default 20)
 *   "tuple_cap_mult", "tuple_cap_extra"   candidate tuple capacity of a find = positions * mult + extra (defaults 6, 2^20);
 *                    on overflow the engine doubles mult and redoes the range (tests force that path)
 *   "rk_restart"     positions before a range from which stage R looks hits up when it tries to restart its
 *                    carried-match machine behind a hit-free stretch instead of at the last ring shift (default 2^18)
 *   "retain"         0 = forget the segments after every find (default 1)
 *   "max_segments"   a range with more retained segments than this behind it is computed from scratch (default 8) */
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
