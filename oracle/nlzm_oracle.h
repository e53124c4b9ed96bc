/* TEST INFRASTRUCTURE — CPU oracle for the NLZM match-finding path.
 *
 * A plain-C restatement of the reference's four match finders (nauful/NLZM 1.03,
 * /root/reference/NLZM.cpp:733-1123) and of the encoder geometry that is part of their
 * semantics (NLZM.cpp:1716-1725, 1782-1798). It is the checker the CUDA engine is compared
 * against; it is NOT a fallback: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. Parity is pinned: tests/test_oracle_vs_ref.py
 * compares it step for step with the reference itself (oracle/_ref/libnlzm_ref.so) and with the
 * fixtures under tests/golden/ that were generated from the reference.
 *
 * Mode: "R2"  — BT4 test cap lifted (optional), skip rule off: every finder is
 * called at every eligible position, so the output is a pure function of (bytes, hist_bits).
 */
#ifndef NLZM_ORACLE_H
#define NLZM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLZM_MATCH_MAX 264u

enum { NLZM_F_HT2 = 1, NLZM_F_HT3 = 2, NLZM_F_BT4 = 4, NLZM_F_RK256 = 8, NLZM_F_ALL = 15 };

typedef struct {
    uint32_t hist_bits;   /* after the shrink loop, NLZM.cpp:1716-1718 */
    uint32_t window;      /* 1 << hist_bits */
    uint32_t frame_bits;  /* NLZM.cpp:1722 */
    uint32_t chunk_size;  /* NLZM.cpp:1724 */
    uint32_t feed_size;   /* NLZM.cpp:1725 */
    uint32_t ht2_bits, ht3_bits, bt4_bits, rk_bits; /* NLZM.cpp:1750-1753 */
} nlzm_geom;

void nlzm_oracle_geometry(uint64_t flen, uint32_t hist_bits_req, nlzm_geom *g);

/* Per-position staircase steps in CSR form: position a owns steps [offsets[a], offsets[a+1]),
 * strictly increasing in both len and dist. */
typedef struct {
    uint64_t n_pos;
    uint64_t n_steps;
    uint64_t *offsets;   /* n_pos + 1 */
    uint32_t *dist;
    uint16_t *len;
} nlzm_steps;

/* bt_max_tests: 0 = unlimited (cap lifted), 256 = as shipped. Returns 0 on success. */
int nlzm_oracle_find(const uint8_t *x, uint64_t flen, uint32_t hist_bits_req, uint32_t finder_mask,
                     uint32_t bt_max_tests, nlzm_steps *out);
void nlzm_oracle_free(nlzm_steps *s);

#ifdef __cplusplus
}
#endif
#endif
