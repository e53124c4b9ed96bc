/* TEST INFRASTRUCTURE — see nlzm_oracle.h. Sequential restatement of the reference matchers. */
#include "nlzm_oracle.h"
#include <stdlib.h>
#include <string.h>

#define NONE 0xFFFFFFFFu
#define HASH_MUL 987660757u          /* NLZM.cpp:739 */
#define RK_BLOCK 256u                /* NLZM.cpp:789-791 */
#define RK_ADDH 0x2F0FD693u          /* NLZM.cpp:793 */
#define RK_REMH 0x0E4EA401u          /* NLZM.cpp:796 = ADDH^256 */

static uint32_t clampu(uint32_t v, uint32_t lo, uint32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

void nlzm_oracle_geometry(uint64_t flen, uint32_t hb, nlzm_geom *g) {
    while (hb > 10 && flen < (1ull << (hb - 1))) --hb;              /* NLZM.cpp:1716-1718 */
    g->hist_bits = hb;
    g->window = 1u << hb;
    g->frame_bits = clampu(hb - 2, 14, 17);                         /* NLZM.cpp:1722 */
    g->chunk_size = ((1u << g->frame_bits) * 15) / 16 - 0x200;      /* NLZM.cpp:1724 */
    g->feed_size = g->chunk_size + NLZM_MATCH_MAX + 1;              /* NLZM.cpp:1725 */
    g->ht2_bits = 12;                                               /* NLZM.cpp:1750 */
    g->ht3_bits = 12 + clampu(hb, 15, 17) - 15;                     /* NLZM.cpp:1751 */
    g->bt4_bits = 13 + clampu(hb, 16, 20) - 16;                     /* NLZM.cpp:1752 */
    g->rk_bits = 15 + clampu(hb, 16, 22) - 16;                      /* NLZM.cpp:1753 */
}

/* distance-dependent minimum match length, NLZM.cpp:813-821 */
static uint32_t match_min(uint32_t d) { return 2 + (d >= 256) + (d >= 4096) + (d >= (1u << 20)); }

/* The per-position candidate staircase, NLZM.cpp:746-752 and Update at 835-852. */
typedef struct { uint32_t max_len; uint32_t delta[NLZM_MATCH_MAX + 2]; } stair;

static void stair_update(stair *t, uint32_t d, uint32_t len) {
    uint32_t i = 0;
    for (; i <= len && i <= t->max_len; i++) if (d < t->delta[i]) t->delta[i] = d;
    for (; i <= len; i++) t->delta[i] = d;
    if (len > t->max_len) t->max_len = len;
}

/* common-prefix length of x[p0..] and x[p1..] (p0 < p1) from `from`, capped; bit 31 of the
 * result = x[p0+m] < x[p1+m] at the first difference (NLZM.cpp:854-877). */
static uint32_t lcp_signed(const uint8_t *x, uint64_t p0, uint64_t p1, uint32_t cap, uint32_t from) {
    uint32_t m = from;
    while (m < cap) {
        uint8_t c0 = x[p0 + m], c1 = x[p1 + m];
        if (c0 != c1) return m | ((uint32_t)(c0 < c1) << 31);
        ++m;
    }
    return m;
}

/* ---- HT2 / HT3: hash rows with overlapping cells (NLZM.cpp:893-957) ---- */
typedef struct { uint32_t bits, rows, nbytes; uint32_t *cells; } ht_t;

static int ht_init(ht_t *h, uint32_t bits, uint32_t rows, uint32_t nbytes) {
    h->bits = bits; h->rows = rows; h->nbytes = nbytes;
    size_t n = (size_t)rows << bits;
    h->cells = (uint32_t *)malloc(n * 4);
    if (!h->cells) return -1;
    memset(h->cells, 0xFF, n * 4);
    return 0;
}

static void ht_find(ht_t *h, stair *t, const uint8_t *x, uint64_t a, uint32_t P, uint64_t shift_total,
                    uint32_t cap, uint32_t hb) {
    uint32_t v = 0;
    for (uint32_t i = 0; i < h->nbytes; i++) v |= (uint32_t)x[a + i] << (8 * i);
    const uint32_t hash = v * HASH_MUL;
    const uint32_t wmask = (1u << hb) - 1;
    const uint32_t chk = hash & ((1u << (32 - hb)) - 1);
    uint32_t *cell = h->cells + (hash >> (32 - h->bits));          /* rows + bucket, NOT bucket*rows (912) */
    uint32_t carry = P | (chk << hb);                              /* full P: may spill into chk (913) */
    uint32_t best = 1;
    for (uint32_t i = 0; i < h->rows; i++) {
        const uint32_t row = cell[i];
        if (best < cap && (row >> hb) == chk) {
            const uint32_t sp = row & wmask;
            if (sp < P && P - sp <= wmask) {
                const uint32_t m = lcp_signed(x, sp + shift_total, a, cap, 0) & 0x7FFFFFFF;
                if (m > best && m >= match_min(P - sp)) { stair_update(t, P - sp, m); best = m; }
            }
        }
        cell[i] = carry;
        carry = row;
    }
}

/* ---- BT4: binary tree per 4-byte-hash bucket (NLZM.cpp:959-1031), absolute coordinates ---- */
typedef struct { uint32_t bits; uint32_t *heads, *tree; } bt_t;

static int bt_init(bt_t *b, uint32_t bits, uint32_t hb) {
    b->bits = bits;
    b->heads = (uint32_t *)malloc(((size_t)4) << bits);
    b->tree = (uint32_t *)malloc(((size_t)8) << hb);
    if (!b->heads || !b->tree) return -1;
    memset(b->heads, 0xFF, ((size_t)4) << bits);
    memset(b->tree, 0xFF, ((size_t)8) << hb);
    return 0;
}

static void bt_find(bt_t *b, stair *t, const uint8_t *x, uint32_t a, uint32_t cap, uint32_t hb,
                    uint32_t max_tests) {
    const uint32_t wmask = (1u << hb) - 1;
    uint32_t v = (uint32_t)x[a] | ((uint32_t)x[a + 1] << 8) | ((uint32_t)x[a + 2] << 16) | ((uint32_t)x[a + 3] << 24);
    const uint32_t bucket = (v * HASH_MUL) >> (32 - b->bits);
    uint32_t *pend_l = b->tree + ((size_t)(a & wmask) << 1);
    uint32_t *pend_r = pend_l + 1;
    uint32_t len_l = 0, len_r = 0;
    uint32_t sp = b->heads[bucket];
    b->heads[bucket] = a;
    uint32_t tests = 0;
    while (sp != NONE && a - sp <= wmask && (max_tests == 0 || tests < max_tests)) {
        ++tests;
        uint32_t *pair = b->tree + ((size_t)(sp & wmask) << 1);
        const uint32_t ms = lcp_signed(x, sp, a, cap, len_l < len_r ? len_l : len_r);
        const uint32_t m = ms & 0x7FFFFFFF;
        if (m >= match_min(a - sp)) stair_update(t, a - sp, m);
        if (m == cap) {                       /* full-length match: adopt its children, drop it */
            *pend_l = pair[0];
            *pend_r = pair[1];
            return;
        }
        if (ms >> 31) {                       /* x[sp..] < x[a..]: sp hangs left, continue right */
            *pend_l = sp; pend_l = pair + 1; sp = *pend_l; len_r = m;
        } else {
            *pend_r = sp; pend_r = pair; sp = *pend_r; len_l = m;
        }
    }
    *pend_r = NONE;
    *pend_l = NONE;
}

/* ---- RK256: rolling hash + one-entry table + carried match (NLZM.cpp:1033-1123) ---- */
typedef struct {
    uint32_t bits; uint32_t *table;
    uint32_t cf, ct, cl;      /* carry_match_from / to / len, shifted coordinates */
    uint32_t h; int primed;
} rk_t;

static int rk_init(rk_t *r, uint32_t bits) {
    r->bits = bits;
    r->table = (uint32_t *)malloc(((size_t)4) << bits);
    if (!r->table) return -1;
    memset(r->table, 0xFF, ((size_t)4) << bits);
    r->cf = r->ct = r->cl = 0; r->h = 0; r->primed = 0;
    return 0;
}

static void rk_find(rk_t *r, stair *t, const uint8_t *x, uint64_t a, uint32_t P, uint64_t shift_total,
                    uint32_t rem, uint32_t hb) {
    const uint32_t wmask = (1u << hb) - 1;
    /* (A) carried match, NLZM.cpp:1056-1069 (u32 wrap-around drops it at a ring shift) */
    if (r->cl > 0) {
        if (P - r->ct < r->cl) {
            const uint32_t d = r->ct - r->cf, m = r->cl - (P - r->ct);
            if (m >= match_min(d)) stair_update(t, d, m < NLZM_MATCH_MAX ? m : NLZM_MATCH_MAX);
        } else r->cl = 0;
    }
    /* (B) hash of x[a .. a+256): rolled one byte per call, NLZM.cpp:1071-1088, 798-799 */
    if (!r->primed) {
        uint32_t h = 0;
        for (uint32_t i = 0; i < RK_BLOCK; i++) h = (x[a + i] + h) * RK_ADDH;
        r->h = h; r->primed = 1;
    } else {
        r->h = ((uint32_t)x[a + RK_BLOCK - 1] + r->h - (uint32_t)x[a - 1] * RK_REMH) * RK_ADDH;
    }
    const uint32_t h = r->h, slot = h >> (32 - r->bits), cmask = (1u << (32 - hb)) - 1;
    /* (C) lookup, NLZM.cpp:1090-1107; length cap is a uint16 in the reference (759-760) */
    if (r->cl < RK_BLOCK) {
        const uint32_t e = r->table[slot];
        const uint32_t sp = e & wmask;
        if ((e >> hb) == (h & cmask) && sp < P && P - sp <= wmask) {
            const uint32_t cap = rem & 0xFFFF;
            const uint32_t m = lcp_signed(x, sp + shift_total, a, cap, 0) & 0x7FFFFFFF;
            if (m >= r->cl && m >= match_min(P - sp)) {
                stair_update(t, P - sp, m < NLZM_MATCH_MAX ? m : NLZM_MATCH_MAX);
                r->cf = sp; r->ct = P; r->cl = m;
            }
        }
    }
    /* (D) insert aligned block after the lookup, NLZM.cpp:1109-1112 */
    if ((P & (RK_BLOCK - 1)) == 0) r->table[slot] = P | (h << hb);
}

/* ---- driver: geometry of encode_file, R2 call pattern of parse_table ---- */
typedef struct { uint64_t cap; nlzm_steps *s; } sink;

static int sink_push(sink *k, uint32_t d, uint32_t len) {
    nlzm_steps *s = k->s;
    if (s->n_steps == k->cap) {
        uint64_t nc = k->cap ? k->cap * 2 : (1u << 16);
        uint32_t *nd = (uint32_t *)realloc(s->dist, nc * 4);
        if (!nd) return -1;
        s->dist = nd;
        uint16_t *nl = (uint16_t *)realloc(s->len, nc * 2);
        if (!nl) return -1;
        s->len = nl;
        k->cap = nc;
    }
    s->dist[s->n_steps] = d;
    s->len[s->n_steps] = (uint16_t)len;
    s->n_steps++;
    return 0;
}

int nlzm_oracle_find(const uint8_t *xin, uint64_t flen, uint32_t hist_bits_req, uint32_t mask,
                     uint32_t bt_max_tests, nlzm_steps *out) {
    memset(out, 0, sizeof *out);
    if (flen >= 0xFFFFFF00ull) return -2;
    nlzm_geom g;
    nlzm_oracle_geometry(flen, hist_bits_req, &g);
    const uint32_t hb = g.hist_bits, W = g.window;
    /* the reference reads 4 bytes at positions whose tail may be short: keep a zero tail */
    uint8_t *x = (uint8_t *)malloc(flen + 16);
    if (!x) return -1;
    memcpy(x, xin, flen);
    memset(x + flen, 0, 16);

    ht_t ht2, ht3; bt_t bt; rk_t rk;
    memset(&ht2, 0, sizeof ht2); memset(&ht3, 0, sizeof ht3); memset(&bt, 0, sizeof bt); memset(&rk, 0, sizeof rk);
    int rc = 0;
    if (ht_init(&ht2, g.ht2_bits, 1, 2) || ht_init(&ht3, g.ht3_bits, 2, 3) || bt_init(&bt, g.bt4_bits, hb) ||
        rk_init(&rk, g.rk_bits)) rc = -1;
    out->n_pos = flen;
    out->offsets = (uint64_t *)malloc((flen + 1) * 8);
    if (!out->offsets) rc = -1;
    sink k = { 0, out };
    stair t;

    uint64_t hist_pos = 0, shift_total = 0, a0 = 0;
    while (rc == 0 && a0 < flen) {
        const uint64_t chunk_read = flen - a0 < g.feed_size ? flen - a0 : g.feed_size;   /* NLZM.cpp:1774,1870-1885 */
        const uint64_t p_end = chunk_read < g.chunk_size ? chunk_read : g.chunk_size;    /* NLZM.cpp:1783 */
        if (hist_pos >= 2ull * W) {                                                      /* NLZM.cpp:1786-1792 */
            hist_pos -= W; shift_total += W;
            ht2.cells[0] = NONE;   /* MatchFinderHT::Shift as written only clears rows[0] (940-957) */
            ht3.cells[0] = NONE;
            /* BT4 is kept in absolute coordinates; RK's table and carry are never shifted (1115-1123) */
        }
        for (uint64_t p = 0; p < p_end && rc == 0; p++) {
            const uint64_t a = a0 + p;
            const uint32_t P = (uint32_t)(hist_pos + p);
            const uint32_t rem = (uint32_t)(chunk_read - p);
            const uint32_t cap = rem < NLZM_MATCH_MAX ? rem : NLZM_MATCH_MAX;            /* NLZM.cpp:915,987 */
            t.max_len = 0;
            out->offsets[a] = out->n_steps;
            if (rem >= 4) {                                                              /* NLZM.cpp:1515 */
                if (mask & NLZM_F_HT2) ht_find(&ht2, &t, x, a, P, shift_total, cap, hb);
                if (mask & NLZM_F_HT3) ht_find(&ht3, &t, x, a, P, shift_total, cap, hb);
                if (mask & NLZM_F_BT4) bt_find(&bt, &t, x, (uint32_t)a, cap, hb, bt_max_tests);
            }
            if (rem >= RK_BLOCK && (mask & NLZM_F_RK256))                                /* NLZM.cpp:1525 */
                rk_find(&rk, &t, x, a, P, shift_total, rem, hb);
            for (uint32_t i = 1; i <= t.max_len; i++)
                if (i == t.max_len || t.delta[i + 1] != t.delta[i])
                    if (sink_push(&k, t.delta[i], i)) { rc = -1; break; }
        }
        hist_pos += p_end;
        a0 += p_end;
    }
    if (rc == 0) out->offsets[flen] = out->n_steps;
    free(ht2.cells); free(ht3.cells); free(bt.heads); free(bt.tree); free(rk.table); free(x);
    if (rc) nlzm_oracle_free(out);
    return rc;
}

void nlzm_oracle_free(nlzm_steps *s) {
    free(s->offsets); free(s->dist); free(s->len);
    memset(s, 0, sizeof *s);
}
