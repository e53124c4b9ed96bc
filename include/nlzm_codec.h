/* nlzm_codec.h — C ABI of the host pipeline around the B200 match-finding engine.
 *
 * The caller side of the nlzm_mf.h boundary, written from scratch: the forward parser that consumes
 * the engine's candidate steps, the adaptive nibble model, the bit + 4-way rANS frame writer, and the
 * matching stream reader. Each entry point names the reference interface it replaces
 * (nauful/NLZM 1.03, NLZM.cpp):
 *
 *   nlzm_codec_compress    <- encode_file(fin, fout, hist_bits)            NLZM.cpp:1711-1910
 *                             = parse_table (1464-1651) + model_encode_*  (1274-1367, 1428-1439)
 *                             + CodeFrame (560-640), with the four finder objects replaced by the
 *                             engine (nlzm_mf.h, double-buffered blocks); the `-window:N` clamp of main()
 *                             (NLZM.cpp:2085) is applied to window_bits
 *   nlzm_codec_decompress  <- decode_file(fin, fout)                       NLZM.cpp:1912-2039
 *   nlzm_codec_free        <- delete[] of the reference's buffers
 *
 * Stream format: identical to the reference's. nlzm_codec_compress emits, byte for byte, the stream
 * the reference encoder emits when its finders are replaced by the engine; the
 * pristine reference decoder restores the input from it, and nlzm_codec_decompress restores the
 * input from streams written by the pristine reference encoder.
 *
 * nlzm_codec_compress needs a CUDA device (it drives libnlzm_mf); there is no CPU fallback for the
 * matcher stage. nlzm_codec_decompress is sequential entropy decoding and runs on the host, as in
 * the reference.
 *
 * Returns 0 or a negative nlzm_codec_status; nlzm_codec_last_error() (thread local) has the text.
 */
#ifndef NLZM_CODEC_H
#define NLZM_CODEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLZM_CODEC_ABI_VERSION 2

enum nlzm_codec_status {
    NLZM_CODEC_OK = 0,
    NLZM_CODEC_E_ARG = -1,      /* bad argument */
    NLZM_CODEC_E_NOMEM = -2,    /* host allocation failed */
    NLZM_CODEC_E_ENGINE = -3,   /* the match-finding engine failed (no device, ...) */
    NLZM_CODEC_E_STREAM = -4    /* decompress: malformed or truncated stream */
};

typedef struct {
    uint32_t struct_size;       /* sizeof(nlzm_codec_config) */
    uint32_t window_bits;       /* -window:N; clamped to 15..28 like the reference CLI (default there: 22) */
    int32_t device;             /* CUDA ordinal for the engine */
    uint32_t reserved;
    uint64_t block_len;         /* positions per engine call; 0 = max(window, 32 Mi), capped at 2^28 */
    /* Several GPUs of one box driven by this process (struct_size tells whether these fields are present):
     * n_devices > 1 replicates the input on devices[0..n_devices) and deals the blocks round-robin; the window
     * behind a block is copied from the engine that owns the block before it (peer copy). `device` is ignored
     * then and blocks are at least one window long. The stream is the same as with one device. */
    uint32_t n_devices;
    int32_t devices[8];
    uint32_t reserved2;
} nlzm_codec_config;

typedef struct {
    uint64_t in_bytes, out_bytes;
    uint64_t literals, matches, reps;     /* commands written */
    uint64_t frames;                      /* one per chunk (NLZM.cpp:1779, 1849) */
    uint64_t parses;                      /* parse_table segments */
    uint64_t steps_served;                /* candidate steps consumed from the engine */
    uint64_t engine_blocks;               /* engine calls fetched */
    double ms_total, ms_engine_wait;      /* wall time; time the parser spent blocked on the engine */
} nlzm_codec_stats;

int nlzm_codec_abi_version(void);

/* *out is allocated by the library (release with nlzm_codec_free). stats may be NULL. */
int nlzm_codec_compress(const uint8_t *in, uint64_t in_len, const nlzm_codec_config *cfg,
                        uint8_t **out, uint64_t *out_len, nlzm_codec_stats *stats);

int nlzm_codec_decompress(const uint8_t *in, uint64_t in_len, uint8_t **out, uint64_t *out_len);

void nlzm_codec_free(uint8_t *p);

const char *nlzm_codec_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
