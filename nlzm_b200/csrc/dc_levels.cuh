// Stage T — exhaustive BT4 by divide and conquer over position ranges.
//
// Reference being replaced: MatchFinderBT::FindAndUpdate, NLZM.cpp:978-1022, with the test cap
// lifted. Its output at position a is, for every length L, the nearest earlier position q in the
// window with lcp(q, a) >= L (verified by brute force). Those (q, a) pairs are
// found level by level: at level k the universe is cut into segments of 2h = 2^k positions, each
// held as an array sorted by suffix rank. For a right-half element a, the left-half candidates
// are the prefix maxima (by position) walking away from a's insertion point in the sorted left
// half — on the rank-left side via "previous greater position" pointers (pg), on the rank-right
// side via "next greater position" pointers (ng). Along such a chain positions get nearer to a
// and lcp(., a) never grows, so the walk stops at the first element that does not beat best[a],
// the longest match already found at nearer levels.
//
// Two kernels do the levels:
//   k_dc_base        one CTA per tile of 2^BASE_LOG positions: text, ranks, level arrays and pointers
//                    live in shared memory for the first BASE_LOG levels
//   k_dc_merge_tile  one level above that: merge-path tiles of 1024 elements staged in shared memory,
//                    coalesced in and out; elements carry their first 22 bytes and their best length,
//                    so most candidate tests never touch the text
//   k_dc_partition   merge-path diagonals for k_dc_merge_tile; k_dc_link: pg/ng maintenance
#pragma once
#include "common.cuh"

#if defined(NLZM_EMU) && defined(NLZM_EMU_STATS)
extern unsigned long long nlzm_stats[16];
#define NLZM_STAT(i, v) __atomic_fetch_add(&nlzm_stats[i], (unsigned long long)(v), __ATOMIC_RELAXED)
#else
#define NLZM_STAT(i, v) ((void)0)
#endif

// Greater-position pointers of one position: 16 bytes, each half = position (universe-relative, low 32 bits) |
// link lcp << 32. No ranks: whether a new right-half neighbour r is rank-nearer to c than the pointer c holds is
// decided by the link lcps alone — lcp(c, r) > lcp(c, pg) means nearer, < means farther, and on a tie r may be taken
// either way: everything between r and pg in rank order then has the same lcp with c, hence with any query that
// reaches c, and r (the later position) is the nearest of them, which is all a staircase keeps.
struct PtrEntry {
    u64 pg;   // nearest element to the rank-left with a greater position in the same segment (NLZM_PTR_NONE if none),
              // link lcp = lcp(this, pg) capped at 264: lcp(a, pg) = min(lcp(a, this), link) along a chain
    u64 ng;   // same to the rank-right
};
#define NLZM_PTR_NONE 0xFFFFFFFFull
HD u64 ptr_pack(u32 pos, u32 lcp) { return (u64)pos | ((u64)lcp << 32); }
HD u32 ptr_pos(u64 e) { return (u32)e; }
HD u32 ptr_lcp(u64 e) { return (u32)(e >> 32) & 0xFFFFu; }

// Level-array element: 32 bytes, streamed once in and once out per level.
struct Elem {
    u64 key;    // rank << 32 | universe-relative position: strict total order of the suffixes
    u64 p0;     // text bytes 0..7 of the suffix (little endian: byte 0 in the low bits)
    u64 p1;     // text bytes 8..15
    u64 tail;   // bits 0..15 text bytes 16..17 | 16..31 best length so far | 32..47 lpg | 48..63 lng
};
#define NLZM_ELEM_PREFIX 18u
HD u32 elem_best(u64 tail) { return (u32)(tail >> 16) & 0xFFFFu; }
HD u32 elem_lpg(u64 tail) { return (u32)(tail >> 32) & 0xFFFFu; }
HD u32 elem_lng(u64 tail) { return (u32)(tail >> 48); }
HD u64 elem_tail(u32 b1617, u32 best, u32 lpg, u32 lng) { return (u64)b1617 | ((u64)best << 16) | ((u64)lpg << 32) | ((u64)lng << 48); }
HD u64 elem_set_best(u64 tail, u32 best) { return (tail & ~0xFFFF0000ull) | ((u64)best << 16); }

struct TupleSink {
    u64 *keys;       // (a_rel << 9) | len
    u32 *vals;       // distance
    u32 *count;      // device counter
    u32 cap;
};

DEV void tuple_append(const TupleSink &s, u32 a_rel, u32 dist, u32 len) {
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__)
    // warp-aggregated append: one atomic per group of converged lanes
    unsigned m = __activemask();
    int leader = __ffs(m) - 1;
    int lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(s.count, (u32)__popc(m));
    base = __shfl_sync(m, base, leader);
    u32 idx = base + __popc(m & ((1u << lane) - 1));
#else
    u32 idx = nlzm_atomic_add(s.count, 1u);
#endif
    if (idx < s.cap) {
        s.keys[idx] = ((u64)a_rel << 9) | len;
        s.vals[idx] = dist;
    }
}

struct DcParams {
    const u8 *x;         // whole input, absolute
    Geom g;
    u64 u0;              // absolute offset of the universe
    u64 own_b, own_e;    // absolute range whose candidates are wanted
    u32 n;               // universe size
    u32 h;               // half segment size at this level (merge levels)
    const u32 *rank;     // stage S output (base kernel only)
    const Elem *cur;     // level k-1 arrays
    Elem *nxt;           // level k arrays
    u32 *corank;         // per cur index of a left-half element: its co-rank in the right half
    u32 origin;          // merge levels: cur/nxt/corank/part are shifted by this many elements (cross pass)
    u32 cross;           // 1 = query-only pass between neighbouring window-sized blocks: nothing is stored
    const u32 *part;     // merge-path split per output tile (merge levels)
    PtrEntry *ptr;       // indexed by universe-relative position
    TupleSink sink;
    u32 n_valid;         // universe-relative positions >= n_valid lie past own_e ("pads", present only so that the
                         // ranks of the last own positions are exact): they sort behind everything and are never adopted
    // cross pass against a RETAINED segment (sorted array + pointers of an earlier find, possibly computed on another
    // GPU, with its own rank universe): the left side is the segment, the right side one sorted block of this find
    const Elem *seg;     // retained segment, sorted by suffix order
    const PtrEntry *seg_ptr;   // its pointers, indexed by position relative to seg_u0
    u64 seg_u0;          // absolute offset its positions are relative to
    u64 seg_last;        // absolute offset of its last position (window pruning)
    u32 seg_len;         // elements in it
    Elem *own;           // right side: a sorted block of this find (best lengths are written back)
    u32 own_len;
};
#define NLZM_RANK_PAD 0xFFFFFFFFu

DEV void dc_emit(const DcParams &p, u64 a_abs, u32 dist, u32 len) {
    if (dist <= p.g.W - 1 && len >= match_min(dist)) tuple_append(p.sink, (u32)(a_abs - p.own_b), dist, len);
}

// can position a_abs be a query at all, and with which length cap
DEV bool dc_query_cap(const DcParams &p, u64 a_abs, u32 &cap) {
    if (a_abs < p.own_b || a_abs >= p.own_e) return false;
    const u64 left_in_file = p.g.flen - a_abs;
    if (left_in_file < 4) return false;                                   // HT/BT need 4 visible bytes (NLZM.cpp:1515)
    cap = left_in_file < NLZM_MATCH_MAX ? (u32)left_in_file : NLZM_MATCH_MAX;   // NLZM.cpp:987
    return true;
}

// ================================================================================================
// base kernel: levels 1..BASE_LOG inside shared memory
// ================================================================================================
#ifndef NLZM_BASE_LOG
#define NLZM_BASE_LOG 12
#endif
#define NLZM_BASE_TILE (1u << NLZM_BASE_LOG)
#ifndef NLZM_BASE_IPT
#define NLZM_BASE_IPT 8u
#endif
#define NLZM_BASE_THREADS (NLZM_BASE_TILE / NLZM_BASE_IPT)
#define NLZM_BASE_TEXT (NLZM_BASE_TILE + NLZM_MATCH_MAX + 8)
#define NLZM_BASE_SMEM (NLZM_BASE_TILE * 20 + NLZM_BASE_TEXT + 8)
#define NLZM_L16_NONE 0xFFFFu

struct BaseSmem {
    u32 *rnk;        // rank of local position
    u32 *k4;         // first 4 text bytes of local position
    u16 *arr[2];     // level arrays: local positions sorted by (rank, position) inside each segment
    u16 *pg, *ng;    // greater-position pointers as local positions
    u16 *best;       // best length per local position
    u16 *cor;        // co-rank in the sibling half, per array index
    u8 *text;        // tile text + lookahead
};

DEV BaseSmem base_carve(u8 *smem) {
    BaseSmem s;
    s.rnk = (u32 *)smem;
    s.k4 = s.rnk + NLZM_BASE_TILE;
    s.arr[0] = (u16 *)(s.k4 + NLZM_BASE_TILE);
    s.arr[1] = s.arr[0] + NLZM_BASE_TILE;
    s.pg = s.arr[1] + NLZM_BASE_TILE;
    s.ng = s.pg + NLZM_BASE_TILE;
    s.best = s.ng + NLZM_BASE_TILE;
    s.cor = s.best + NLZM_BASE_TILE;
    s.text = (u8 *)(s.cor + NLZM_BASE_TILE);
    return s;
}

DEV bool base_less(const BaseSmem &s, u32 a, u32 b) {       // (rank, position) order; ranks tie only for equal prefixes
    const u32 ra = s.rnk[a], rb = s.rnk[b];
    return ra < rb || (ra == rb && a < b);
}

DEV u32 base_lower_bound(const BaseSmem &s, const u16 *arr, u32 n, u32 pos) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (base_less(s, arr[mid], pos)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// merge-path split inside shared memory: left elements among the first d outputs of the segment
DEV u32 base_merge_path(const BaseSmem &s, const u16 *L, u32 l_len, const u16 *R, u32 r_len, u32 d) {
    u32 lo = d > r_len ? d - r_len : 0, hi = d < l_len ? d : l_len;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (base_less(s, L[mid], R[d - 1 - mid])) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Walk one chain; the caller has already established that the first 4 bytes of `start` match.
DEV u32 base_lcp4(const BaseSmem &s, u32 c, u32 a, u32 lim) {     // first 4 bytes are known to be equal
    u32 l = 4;
    while (l < lim && s.text[c + l] == s.text[a + l]) ++l;
    return l < lim ? l : lim;
}

// (dom_c, dom_l): the other side's first candidate, see dc_walk
DEV void base_walk(const DcParams &p, const BaseSmem &s, u32 a, u64 a_abs, u32 best_in, u32 start, u32 first_l, const u16 *link,
                   u32 dom_c, u32 dom_l, u32 &new_best) {
    u32 c = start, l = first_l, pend_len = 0, pend_c = 0;
    const u32 ka = s.k4[a];
    while (l > best_in) {
        if (pend_len && l < pend_len && !(dom_l >= pend_len && dom_c > pend_c)) dc_emit(p, a_abs, a - pend_c, pend_len);
        pend_len = l; pend_c = c;       // equal lcp: the nearer element replaces the farther one
        if (l > new_best) new_best = l;
        c = link[c];
        if (c == NLZM_L16_NONE || s.k4[c] != ka) break;      // lcp < 4 can never beat best (>= 3)
        l = base_lcp4(s, c, a, l);                           // lcp never grows along the chain
    }
    if (pend_len && !(dom_l >= pend_len && dom_c > pend_c)) dc_emit(p, a_abs, a - pend_c, pend_len);
}

DEV void dc_base_cta(const DcParams &p, u32 bid, u32 tid, u8 *smem) {
    const BaseSmem s = base_carve(smem);
    const u32 t0 = bid * NLZM_BASE_TILE;
    const u32 m = (p.n - t0) < NLZM_BASE_TILE ? (p.n - t0) : NLZM_BASE_TILE;
    const u64 abs0 = p.u0 + t0;
    const u64 x_end = p.g.flen + NLZM_X_PAD;                 // bytes [flen, x_end) are zero padding
    // tile text: 8 bytes per load (the tile may start at any byte offset: two aligned words and a funnel shift)
    for (u32 i = tid * 8; i < NLZM_BASE_TEXT; i += NLZM_BASE_THREADS * 8) {
        u64 w = 0;
        if (abs0 + i + 8 <= x_end) w = load8(p.x, abs0 + i);
        else for (u32 k = 0; k < 8; k++) if (abs0 + i + k < x_end) w |= (u64)p.x[abs0 + i + k] << (8 * k);
        if (i + 8 <= NLZM_BASE_TEXT) *(u64 *)(s.text + i) = w;
        else for (u32 k = 0; i + k < NLZM_BASE_TEXT; k++) s.text[i + k] = (u8)(w >> (8 * k));
    }
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__)
    {
        // ranks of a full tile: one bulk copy through the TMA unit (16 KiB, contiguous, 16-byte aligned)
        __shared__ __align__(8) u64 rank_bar;
        const u32 bar = nlzm_smem_addr(&rank_bar);
        const bool bulk = m == NLZM_BASE_TILE && (((uintptr_t)(p.rank + t0)) & 15) == 0;
        if (bulk) {
            if (tid == 0) nlzm_mbar_init(bar, 1);
            __syncthreads();
            if (tid == 0) {
                nlzm_mbar_expect_tx(bar, NLZM_BASE_TILE * 4);
                nlzm_bulk_g2s(nlzm_smem_addr(s.rnk), p.rank + t0, NLZM_BASE_TILE * 4, bar);
            }
            nlzm_mbar_wait(bar, 0);
        }
        for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
            if (!bulk) s.rnk[i] = p.rank[t0 + i];
            if ((t0 + i) >= p.n_valid) s.rnk[i] = NLZM_RANK_PAD;
            s.arr[0][i] = (u16)i;
            s.pg[i] = NLZM_L16_NONE;
            s.ng[i] = NLZM_L16_NONE;
            s.best[i] = 3;
        }
    }
#else
    for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
        s.rnk[i] = (t0 + i) < p.n_valid ? p.rank[t0 + i] : NLZM_RANK_PAD;
        s.arr[0][i] = (u16)i;
        s.pg[i] = NLZM_L16_NONE;
        s.ng[i] = NLZM_L16_NONE;
        s.best[i] = 3;
    }
#endif
    NLZM_CTA_SYNC();
    for (u32 i = tid; i < m; i += NLZM_BASE_THREADS)
        s.k4[i] = (u32)s.text[i] | ((u32)s.text[i + 1] << 8) | ((u32)s.text[i + 2] << 16) | ((u32)s.text[i + 3] << 24);
    NLZM_CTA_SYNC();
    u32 cur = 0;
    for (u32 h = 1; h < m; h <<= 1) {
        const u16 *A = s.arr[cur];
        u16 *B = s.arr[cur ^ 1];
        // ---- phase A: merge by co-rank; cor[i] = co-rank of A[i] in the sibling half
        if (2 * h < NLZM_BASE_IPT) {
            for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
                const u32 pos = A[i];
                const u32 base = i & ~(2 * h - 1);
                const u32 l_len = (m - base) < h ? (m - base) : h;
                const u32 r_beg = base + l_len;
                const u32 r_len = (m - r_beg) < h ? (m - r_beg) : h;
                if (r_len == 0) { B[i] = (u16)pos; continue; }
                if (i < r_beg) {
                    const u32 u = base_lower_bound(s, A + r_beg, r_len, pos);
                    s.cor[i] = (u16)u;
                    B[i + u] = (u16)pos;
                } else {
                    const u32 t = base_lower_bound(s, A + base, l_len, pos);
                    s.cor[i] = (u16)t;
                    B[base + (i - r_beg) + t] = (u16)pos;
                }
            }
        } else {
            // each thread produces NLZM_BASE_IPT consecutive outputs of one segment (IPT divides 2h)
            const u32 o_beg = tid * NLZM_BASE_IPT;
            if (o_beg < m) {
                const u32 base = o_beg & ~(2 * h - 1);
                const u32 l_len = (m - base) < h ? (m - base) : h;
                const u32 r_beg = base + l_len;
                const u32 r_len = (m - r_beg) < h ? (m - r_beg) : h;
                const u32 o_end = (o_beg + NLZM_BASE_IPT) < (r_beg + r_len) ? (o_beg + NLZM_BASE_IPT) : (r_beg + r_len);
                if (r_len == 0) {
                    for (u32 o = o_beg; o < o_end; o++) B[o] = A[o];
                } else {
                    u32 li = base_merge_path(s, A + base, l_len, A + r_beg, r_len, o_beg - base);
                    u32 ri = (o_beg - base) - li;
                    u32 pl = li < l_len ? A[base + li] : 0, pr = ri < r_len ? A[r_beg + ri] : 0;
                    u32 rl = li < l_len ? s.rnk[pl] : 0, rr = ri < r_len ? s.rnk[pr] : 0;
                    for (u32 o = o_beg; o < o_end; o++) {
                        const bool take_l = (ri >= r_len) || (li < l_len && (rl < rr || (rl == rr && pl < pr)));
                        if (take_l) {
                            B[o] = (u16)pl;
                            s.cor[base + li] = (u16)ri;
                            ++li;
                            if (li < l_len) { pl = A[base + li]; rl = s.rnk[pl]; }
                        } else {
                            B[o] = (u16)pr;
                            s.cor[r_beg + ri] = (u16)li;
                            ++ri;
                            if (ri < r_len) { pr = A[r_beg + ri]; rr = s.rnk[pr]; }
                        }
                    }
                }
            }
        }
        NLZM_CTA_SYNC();
        // ---- phase Q: right-half elements query the left half
        for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
            const u32 base = i & ~(2 * h - 1);
            const u32 l_len = (m - base) < h ? (m - base) : h;
            const u32 r_beg = base + l_len;
            if (i < r_beg || r_beg >= m) continue;
            const u32 pos = A[i], t = s.cor[i];
            const u32 cl = t > 0 ? A[base + t - 1] : NLZM_L16_NONE;
            const u32 cr = t < l_len ? A[base + t] : NLZM_L16_NONE;
            const u32 ka = s.k4[pos];
            const bool hit_l = cl != NLZM_L16_NONE && s.k4[cl] == ka;
            const bool hit_r = cr != NLZM_L16_NONE && s.k4[cr] == ka;
            if (!hit_l && !hit_r) continue;
            u32 cap;
            const u64 a_abs = abs0 + pos;
            if (!dc_query_cap(p, a_abs, cap)) continue;
            const u32 best_in = s.best[pos];
            if (best_in >= cap) continue;
            u32 nb = best_in;
            const u32 ll = hit_l ? base_lcp4(s, cl, pos, cap) : 0u, lr = hit_r ? base_lcp4(s, cr, pos, cap) : 0u;
            if (hit_l) base_walk(p, s, pos, a_abs, best_in, cl, ll, s.pg, cr, lr > best_in ? lr : 0u, nb);
            if (hit_r) base_walk(p, s, pos, a_abs, best_in, cr, lr, s.ng, cl, ll > best_in ? ll : 0u, nb);
            if (nb != best_in) s.best[pos] = (u16)nb;
        }
        NLZM_CTA_SYNC();
        // ---- phase B: left-half elements adopt their rank-nearest right neighbours when nearer in rank
        for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
            const u32 base = i & ~(2 * h - 1);
            const u32 l_len = (m - base) < h ? (m - base) : h;
            const u32 r_beg = base + l_len;
            if (i >= r_beg) continue;
            const u32 r_len = (m - r_beg) < h ? (m - r_beg) : h;
            if (r_len == 0) continue;
            const u32 pos = A[i], u = s.cor[i];
            if (u > 0) { const u32 r = A[r_beg + u - 1]; const u32 o = s.pg[pos]; if (o == NLZM_L16_NONE || base_less(s, o, r)) s.pg[pos] = (u16)r; }
            if (u < r_len) { const u32 r = A[r_beg + u]; const u32 o = s.ng[pos]; if (s.rnk[r] != NLZM_RANK_PAD && (o == NLZM_L16_NONE || base_less(s, r, o))) s.ng[pos] = (u16)r; }
        }
        NLZM_CTA_SYNC();
        cur ^= 1;
    }
    // link lcps of the final pointers (reuse best/cor storage is not possible: keep them in registers via smem cor/arr[cur^1])
    u16 *LPG = s.arr[cur ^ 1];      // the other level array is free now
    u16 *LNG = s.cor;
    for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
        const u32 g0 = s.pg[i], g1 = s.ng[i];
        u32 l0 = 0, l1 = 0;
        if (g0 != NLZM_L16_NONE && s.k4[g0] == s.k4[i]) {
            const u64 later = abs0 + (g0 > i ? g0 : i);
            const u32 lim = (p.g.flen - later) < NLZM_MATCH_MAX ? (u32)(p.g.flen - later) : NLZM_MATCH_MAX;
            l0 = 4; while (l0 < lim && s.text[g0 + l0] == s.text[i + l0]) ++l0;
            if (l0 > lim) l0 = lim;
        } else if (g0 != NLZM_L16_NONE) {
            const u32 d = s.k4[g0] ^ s.k4[i];
            l0 = (u32)nlzm_ctz64((u64)d | (1ull << 32)) >> 3;
        }
        if (g1 != NLZM_L16_NONE && s.k4[g1] == s.k4[i]) {
            const u64 later = abs0 + (g1 > i ? g1 : i);
            const u32 lim = (p.g.flen - later) < NLZM_MATCH_MAX ? (u32)(p.g.flen - later) : NLZM_MATCH_MAX;
            l1 = 4; while (l1 < lim && s.text[g1 + l1] == s.text[i + l1]) ++l1;
            if (l1 > lim) l1 = lim;
        } else if (g1 != NLZM_L16_NONE) {
            const u32 d = s.k4[g1] ^ s.k4[i];
            l1 = (u32)nlzm_ctz64((u64)d | (1ull << 32)) >> 3;
        }
        LPG[i] = (u16)l0;
        LNG[i] = (u16)l1;
    }
    NLZM_CTA_SYNC();
    // hand over to the merge levels: fat elements in rank order, pointers by position
    const u16 *A = s.arr[cur];
    for (u32 i = tid; i < m; i += NLZM_BASE_THREADS) {
        const u32 pos = A[i];
        const u8 *t = s.text + pos;
        Elem e;
        e.key = ((u64)s.rnk[pos] << 32) | (t0 + pos);
        u64 a = 0, b = 0;
        #pragma unroll
        for (int k = 0; k < 8; k++) { a |= (u64)t[k] << (8 * k); b |= (u64)t[8 + k] << (8 * k); }
        e.p0 = a; e.p1 = b;
        e.tail = elem_tail((u32)t[16] | ((u32)t[17] << 8), s.best[pos], LPG[pos], LNG[pos]);
        p.nxt[t0 + i] = e;
        PtrEntry pe;
        const u32 g0 = s.pg[i], g1 = s.ng[i];
        pe.pg = g0 == NLZM_L16_NONE ? NLZM_PTR_NONE : ptr_pack(t0 + g0, LPG[i]);
        pe.ng = g1 == NLZM_L16_NONE ? NLZM_PTR_NONE : ptr_pack(t0 + g1, LNG[i]);
        p.ptr[t0 + i] = pe;
    }
}
#ifndef NLZM_BASE_MINBLOCKS
#define NLZM_BASE_MINBLOCKS 1
#endif
NLZM_KERNEL_CTA_OCC(dc_base, DcParams, NLZM_BASE_THREADS, NLZM_BASE_MINBLOCKS)

// ================================================================================================
// merge levels
// ================================================================================================
#ifndef NLZM_MT_THREADS
#define NLZM_MT_THREADS 256
#endif
#ifndef NLZM_MT_TMA
#define NLZM_MT_TMA 1                    // 1: tiles staged by TMA bulk copies, 0: by 16-byte loads of all threads
#endif
#ifndef NLZM_MT_ITEMS
#define NLZM_MT_ITEMS 4
#endif
#define NLZM_MT_TILE (NLZM_MT_THREADS * NLZM_MT_ITEMS)
#define NLZM_MT_SMEM ((NLZM_MT_TILE + 2) * 32 + NLZM_MT_TILE * 8 + 16)

struct alignas(16) V16 { u64 a, b; };

struct SegGeom { u32 base, l_len, r_beg, r_len; };
DEV SegGeom seg_geom(u32 n, u32 h, u32 idx) {
    SegGeom s;
    s.base = (idx / (2 * h)) * (2 * h);          // 2h <= 2^31 guaranteed by the host loop (h < n < 2^31)
    s.l_len = (n - s.base) < h ? (n - s.base) : h;
    s.r_beg = s.base + s.l_len;
    s.r_len = (n - s.r_beg) < h ? (n - s.r_beg) : h;
    return s;
}

// merge-path split: number of left elements among the first d merged outputs of a segment
DEV u32 merge_path(const Elem *L, u32 l_len, const Elem *R, u32 r_len, u32 d) {
    u32 lo = d > r_len ? d - r_len : 0, hi = d < l_len ? d : l_len;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (L[mid].key < R[d - 1 - mid].key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one thread per output tile boundary
DEV void dc_partition_body(const DcParams &p, u64 j) {
    const u32 o = (u32)j * NLZM_MT_TILE;
    const SegGeom s = seg_geom(p.n, p.h, o);
    u32 *part = (u32 *)p.part;
    part[j] = s.r_len ? merge_path(p.cur + s.base, s.l_len, p.cur + s.r_beg, s.r_len, o - s.base) : (o - s.base);
}
NLZM_KERNEL_1D(dc_partition, DcParams)

// lcp of the suffixes behind two elements from their 22-byte prefixes (22 = "at least 22")
DEV u32 elem_prefix_lcp(const Elem &a, const Elem &b) {
    u64 d = a.p0 ^ b.p0;
    if (d) return (u32)nlzm_ctz64(d) >> 3;
    d = a.p1 ^ b.p1;
    if (d) return 8 + ((u32)nlzm_ctz64(d) >> 3);
    d = (a.tail ^ b.tail) & 0xFFFFull;
    if (d) return 16 + ((u32)nlzm_ctz64(d) >> 3);
    return NLZM_ELEM_PREFIX;
}

// full lcp of the suffixes behind two elements (prefixes first, text beyond), capped at lim
DEV u32 elem_pair_lcp2(const DcParams &p, const Elem &a, u64 a_u0, const Elem &b, u64 b_u0, u32 lim) {
    u32 l = elem_prefix_lcp(a, b);
    if (l >= NLZM_ELEM_PREFIX && lim > NLZM_ELEM_PREFIX)
        l = NLZM_ELEM_PREFIX + lcp_cap(p.x, a_u0 + (u32)a.key + NLZM_ELEM_PREFIX, b_u0 + (u32)b.key + NLZM_ELEM_PREFIX, lim - NLZM_ELEM_PREFIX);
    return l < lim ? l : lim;
}
DEV u32 elem_pair_lcp(const DcParams &p, const Elem &a, const Elem &b, u32 lim) { return elem_pair_lcp2(p, a, p.u0, b, p.u0, lim); }

// Walk one greater-position chain of the left half. The first element is a rank neighbour whose element
// (prefix, link lcp) sits in shared memory; the lcp with every further chain element follows from the
// link lcps stored with the pointers: lcp(a, pg(c)) = min(lcp(a, c), lcp(c, pg(c))). One 32-byte ptr[]
// read per hop, none at all when the chain ends at the neighbour.
// (dom_c, dom_l): the other side's first candidate; anything it dominates (nearer and at least as long) would
// only be merged away later, so it is not queued.
// (ptr, delta): the pointer table of the chain's side and how far its position origin lies before p.u0 (0 inside one find)
DEV void dc_walk(const DcParams &p, const Elem &ea, u64 a_abs, u32 best_in, const Elem *first, u32 first_l, bool left,
                 u32 dom_c, u32 dom_l, u32 &new_best, const PtrEntry *__restrict__ ptr, u32 delta) {
    if (!first) return;
    const u32 a_rel = (u32)ea.key + delta;
    u32 c = (u32)first->key;
    u32 l = first_l;
    u32 link = left ? elem_lpg(first->tail) : elem_lng(first->tail);
    bool have_entry = false;
    u64 en = 0;
    u32 pend_len = 0, pend_c = 0;
    while (l > best_in) {
        if (pend_len && l < pend_len && !(dom_l >= pend_len && dom_c > pend_c)) dc_emit(p, a_abs, a_rel - pend_c, pend_len);
        pend_len = l; pend_c = c;                              // equal lcp: the nearer element replaces the farther one
        if (l > new_best) new_best = l;
        const u32 l_next = l < link ? l : link;                // lcp never grows along the chain
        if (l_next <= best_in) break;                          // the chain ends here without touching memory
        if (!have_entry) en = left ? ptr[c].pg : ptr[c].ng;    // the neighbour's pointer (its link lcp came with the element)
        if (ptr_pos(en) == (u32)NLZM_PTR_NONE) break;
        c = ptr_pos(en);
        l = l_next;
        en = left ? ptr[c].pg : ptr[c].ng;                     // 8 bytes: where c points and the lcp of that link
        have_entry = true;
        link = ptr_lcp(en);
    }
    if (pend_len && !(dom_l >= pend_len && dom_c > pend_c)) dc_emit(p, a_abs, a_rel - pend_c, pend_len);
}

// One CTA produces NLZM_MT_TILE consecutive elements of the merged level array:
//   load (coalesced) -> merge path on keys -> neighbour tests on the carried prefixes -> the few
//   elements that found something walk their chains (compacted, all lanes busy) -> store (coalesced)
DEV void dc_merge_tile_cta(const DcParams &p, u32 bid, u32 tid, u8 *smem) {
    Elem *el = (Elem *)smem;                                    // [0] left halo, [1..nl] left part, [nl+1] halo, then right part
    u16 *src = (u16 *)(smem + (NLZM_MT_TILE + 2) * 32);         // output -> slot in el[]
    u16 *lcnt = src + NLZM_MT_TILE;                             // output (right elements) -> left elements before it
    u16 *pos_l = lcnt + NLZM_MT_TILE;                           // left element -> output
    u16 *act = pos_l + NLZM_MT_TILE;                            // outputs that need a chain walk
    u32 *act_n = (u32 *)(act + NLZM_MT_TILE);

    const u32 o0 = bid * NLZM_MT_TILE;
    const SegGeom s = seg_geom(p.n, p.h, o0);
    const u32 seg_end = s.r_beg + s.r_len;
    const u32 o1 = (o0 + NLZM_MT_TILE) < seg_end ? (o0 + NLZM_MT_TILE) : seg_end;
    const u32 cnt = o1 - o0;
    if (s.r_len == 0) {                                         // lonely left half at the end of the universe
        if (!p.cross) for (u32 i = tid; i < cnt; i += NLZM_MT_THREADS) p.nxt[o0 + i] = p.cur[o0 + i];
        return;
    }
    const u32 d0 = o0 - s.base, d1 = o1 - s.base;
    const u32 l0 = p.part[bid];
    const u32 l1 = (o1 < seg_end) ? p.part[bid + 1] : s.l_len;
    const u32 r0 = d0 - l0, r1 = d1 - l1;
    const u32 nl = l1 - l0, nr = r1 - r0;
    Elem *SL = el + 1;
    Elem *SR = el + nl + 2;
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__) && NLZM_MT_TMA
    {
        // The tile's two source ranges are contiguous in the level array: the left part with one halo element on each
        // side, and the right part. One thread hands both to the TMA unit as bulk copies (no LSU traffic, no address
        // arithmetic in the other 255 threads); completion is counted in bytes on an mbarrier.
        __shared__ __align__(8) u64 tile_bar;
        const u32 bar = nlzm_smem_addr(&tile_bar);
        const u32 g_lo = l0 > 0 ? l0 - 1 : 0;                                    // first left element copied
        const u32 g_hi = (l0 + nl + 1) < s.l_len ? (l0 + nl + 1) : s.l_len;      // one past the last
        if (tid == 0) nlzm_mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            const u32 bytes_l = (g_hi - g_lo) * (u32)sizeof(Elem), bytes_r = nr * (u32)sizeof(Elem);
            nlzm_mbar_expect_tx(bar, bytes_l + bytes_r);
            if (bytes_l) nlzm_bulk_g2s(nlzm_smem_addr(el + (g_lo + 1 - l0)), p.cur + s.base + g_lo, bytes_l, bar);
            if (bytes_r) nlzm_bulk_g2s(nlzm_smem_addr(SR), p.cur + s.r_beg + r0, bytes_r, bar);
            *act_n = 0;
        }
        if (tid < 32) nlzm_mbar_wait(bar, 0);          // one warp polls; the others sleep in the barrier below
    }
#else
    {
        // coalesced 16-byte loads: left part with one halo element on each side, then the right part
        const V16 *GL = (const V16 *)(p.cur + s.base), *GR = (const V16 *)(p.cur + s.r_beg);
        V16 *S = (V16 *)el;
        for (u32 c = tid; c < (nl + 2) * 2; c += NLZM_MT_THREADS) {
            const i64 gi = (i64)l0 + (i64)(c >> 1) - 1;
            if (gi >= 0 && gi < (i64)s.l_len) S[c] = GL[gi * 2 + (c & 1)];
        }
        V16 *S2 = (V16 *)SR;
        for (u32 c = tid; c < nr * 2; c += NLZM_MT_THREADS) S2[c] = GR[(u64)r0 * 2 + c];
    }
    if (tid == 0) *act_n = 0;
#endif
    NLZM_CTA_SYNC();

    // ---- merge path on keys: NLZM_MT_ITEMS outputs per thread
    {
        const u32 t_d = tid * NLZM_MT_ITEMS < cnt ? tid * NLZM_MT_ITEMS : cnt;
        const u32 t_e = t_d + NLZM_MT_ITEMS < cnt ? t_d + NLZM_MT_ITEMS : cnt;
        u32 li = merge_path(SL, nl, SR, nr, t_d), ri = t_d - li;
        u64 kl = li < nl ? SL[li].key : 0, kr = ri < nr ? SR[ri].key : 0;
        for (u32 o = t_d; o < t_e; o++) {
            const bool take_l = (ri >= nr) || (li < nl && kl < kr);
            if (take_l) {
                src[o] = (u16)(1 + li);
                pos_l[li] = (u16)o;
                ++li;
                if (li < nl) kl = SL[li].key;
            } else {
                src[o] = (u16)(nl + 2 + ri);
                lcnt[o] = (u16)li;
                ++ri;
                if (ri < nr) kr = SR[ri].key;
            }
        }
    }
    NLZM_CTA_SYNC();

    // ---- neighbour tests from the carried prefixes; survivors go to the active list
    for (u32 o = tid; o < cnt; o += NLZM_MT_THREADS) {
        const u32 slot = src[o];
        if (slot < nl + 2) continue;                            // left element: nothing to query
        const Elem &e = el[slot];
        const u32 pos = (u32)e.key;
        const u64 a_abs = p.u0 + pos;
        u32 cap;
        const u32 best_in = elem_best(e.tail);
        if (!dc_query_cap(p, a_abs, cap) || best_in >= cap || pos - (p.origin + s.r_beg - 1) > p.g.W - 1) continue;
        const u32 li = lcnt[o];
        NLZM_STAT(0, 1);
        bool want = false;
        if (l0 + li > 0) { u32 l = elem_prefix_lcp(e, el[li]); l = l < cap ? l : cap; want |= (l > best_in) || (l >= NLZM_ELEM_PREFIX && cap > NLZM_ELEM_PREFIX); }
        if (l0 + li < s.l_len) { u32 l = elem_prefix_lcp(e, el[li + 1]); l = l < cap ? l : cap; want |= (l > best_in) || (l >= NLZM_ELEM_PREFIX && cap > NLZM_ELEM_PREFIX); }
        if (want) { NLZM_STAT(1, 1); act[nlzm_atomic_add(act_n, 1u)] = (u16)o; }
    }
    NLZM_CTA_SYNC();

    // ---- chain walks of the active elements
    const u32 n_act = *act_n;
    for (u32 k = tid; k < n_act; k += NLZM_MT_THREADS) {
        const u32 o = act[k];
        Elem &e = el[src[o]];
        const Elem ea = e;
        const u32 pos = (u32)ea.key;
        const u64 a_abs = p.u0 + pos;
        u32 cap = 0;
        dc_query_cap(p, a_abs, cap);
        const u32 best_in = elem_best(ea.tail);
        const u32 li = lcnt[o];
        u32 nb = best_in;
        const Elem *fl = (l0 + li > 0) ? &el[li] : nullptr, *fr = (l0 + li < s.l_len) ? &el[li + 1] : nullptr;
        const u32 ll = fl ? elem_pair_lcp(p, ea, *fl, cap) : 0u, lr = fr ? elem_pair_lcp(p, ea, *fr, cap) : 0u;
        const u32 cl = fl ? (u32)fl->key : 0u, cr = fr ? (u32)fr->key : 0u;
        dc_walk(p, ea, a_abs, best_in, fl, ll, true, cr, lr > best_in ? lr : 0u, nb, p.ptr, 0u);
        dc_walk(p, ea, a_abs, best_in, fr, lr, false, cl, ll > best_in ? ll : 0u, nb, p.ptr, 0u);
        if (nb != best_in) {
            e.tail = elem_set_best(ea.tail, nb);
            // query-only pass: nothing is stored, but the next pass over the same elements should know this length
            if (p.cross) ((Elem *)p.cur)[s.r_beg + r0 + (src[o] - (nl + 2))].tail = e.tail;
        }
    }
    NLZM_CTA_SYNC();

    // ---- coalesced stores: merged elements, and the co-rank of every left element
    if (!p.cross) {
        const V16 *S = (const V16 *)el;
        V16 *G = (V16 *)(p.nxt + o0);
        for (u32 c = tid; c < cnt * 2; c += NLZM_MT_THREADS) G[c] = S[(u32)src[c >> 1] * 2 + (c & 1)];
        for (u32 i = tid; i < nl; i += NLZM_MT_THREADS) p.corank[s.base + l0 + i] = r0 + ((u32)pos_l[i] - i);
    }
}
#ifndef NLZM_MT_MINBLOCKS
#define NLZM_MT_MINBLOCKS 5
#endif
NLZM_KERNEL_CTA_OCC(dc_merge_tile, DcParams, NLZM_MT_THREADS, NLZM_MT_MINBLOCKS)

// Left-half elements adopt their rank-nearest right-half neighbours as pg / ng when those are nearer
// in rank than the pointers they already hold (every right-half position is greater). Separate launch:
// the queries of a level must read the pointers of the level below.
DEV void dc_link_body(const DcParams &p, u64 idx64) {
    const u32 idx = (u32)idx64;
    const SegGeom s = seg_geom(p.n, p.h, idx);
    if (idx >= s.r_beg || s.r_len == 0) return;
    const Elem e = p.cur[idx];
    const u32 pos = (u32)e.key;
    const u32 u = p.corank[idx];
    // the link lcps travel with the element, so nothing is read from ptr[]: a side that changes is one 8-byte store
    u32 lpg = elem_lpg(e.tail), lng = elem_lng(e.tail);
    bool ch = false;
    if (u > 0) {
        const Elem r = p.cur[s.r_beg + u - 1];
        const u64 left_in_file = p.g.flen - (p.u0 + (u32)r.key);              // r is the later position
        const u32 l = elem_pair_lcp(p, e, r, left_in_file < NLZM_MATCH_MAX ? (u32)left_in_file : NLZM_MATCH_MAX);
        if (l >= lpg) { p.ptr[pos].pg = ptr_pack((u32)r.key, l); lpg = l; ch = true; }
    }
    if (u < s.r_len) {
        const Elem r = p.cur[s.r_beg + u];
        if ((u32)(r.key >> 32) != NLZM_RANK_PAD) {
            const u64 left_in_file = p.g.flen - (p.u0 + (u32)r.key);
            const u32 l = elem_pair_lcp(p, e, r, left_in_file < NLZM_MATCH_MAX ? (u32)left_in_file : NLZM_MATCH_MAX);
            if (l >= lng) { p.ptr[pos].ng = ptr_pack((u32)r.key, l); lng = l; ch = true; }
        }
    }
    // the merged copy of this element (already written by k_dc_merge_tile) carries the link lcps too
    if (ch) p.nxt[idx + u].tail = elem_tail((u32)e.tail & 0xFFFFu, elem_best(e.tail), lpg, lng);
}
NLZM_KERNEL_1D(dc_link, DcParams)

// ================================================================================================
// cross pass against a retained segment
// ================================================================================================
// A find keeps its final sorted blocks and pointers ("segments"); a later find whose range follows does not
// re-rank and re-merge the window behind it but queries those segments. Ranks of different finds (or GPUs) are
// not comparable, so the merge order comes from the suffixes themselves: the carried 18-byte prefixes first, the
// text beyond them when those tie. Both sides are sorted by an order that refines "first 264 bytes, zero-padded
// past the end of the file", so merging them under that order is well defined; ties put the (earlier) segment side
// first, which is what the rank order does with equal ranks.

// does segment element l go before own element r? Order of the first 264 bytes with the text zero-padded past
// its end — the preorder both rank orders refine (prefix doubling compares zero-padded 8-byte blocks; a block
// that starts past the end only sorts below a block of real zeros, which this order leaves as a tie).
DEV bool x_left_first(const DcParams &p, const Elem &l, const Elem &r) {
    u64 d = l.p0 ^ r.p0;
    if (d) { const u32 sh = (u32)nlzm_ctz64(d) & ~7u; return ((l.p0 >> sh) & 0xFF) < ((r.p0 >> sh) & 0xFF); }
    d = l.p1 ^ r.p1;
    if (d) { const u32 sh = (u32)nlzm_ctz64(d) & ~7u; return ((l.p1 >> sh) & 0xFF) < ((r.p1 >> sh) & 0xFF); }
    d = (l.tail ^ r.tail) & 0xFFFFull;
    if (d) { const u32 sh = (u32)nlzm_ctz64(d) & ~7u; return ((l.tail >> sh) & 0xFF) < ((r.tail >> sh) & 0xFF); }
    // 18 equal bytes (the carried prefixes are zero-padded already). r is the later position.
    const u64 pl = p.seg_u0 + (u32)l.key, pr = p.u0 + (u32)r.key;
    const u64 left_r = p.g.flen - pr, left_l = p.g.flen - pl;
    u32 m = NLZM_ELEM_PREFIX;
    const u32 lim = left_r < NLZM_MATCH_MAX ? (u32)left_r : NLZM_MATCH_MAX;       // bytes r really has
    if (lim > NLZM_ELEM_PREFIX) {
        m += lcp_cap(p.x, pl + NLZM_ELEM_PREFIX, pr + NLZM_ELEM_PREFIX, lim - NLZM_ELEM_PREFIX);
        if (m < lim) return p.x[pl + m] < p.x[pr + m];
    } else {
        m = lim;                                                                 // r ends inside the prefix
    }
    // r is all padding from here on: l goes first only if it is zero up to the full depth as well
    const u32 end_l = left_l < NLZM_MATCH_MAX ? (u32)left_l : NLZM_MATCH_MAX;
    for (; m < end_l; m++) if (p.x[pl + m]) return false;
    return true;
}

DEV u32 x_merge_path(const DcParams &p, const Elem *L, u32 l_len, const Elem *R, u32 r_len, u32 d) {
    u32 lo = d > r_len ? d - r_len : 0, hi = d < l_len ? d : l_len;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (x_left_first(p, L[mid], R[d - 1 - mid])) lo = mid + 1; else hi = mid;
    }
    return lo;
}

DEV void dc_xpartition_body(const DcParams &p, u64 j) {
    const u32 o = (u32)j * NLZM_MT_TILE;
    ((u32 *)p.part)[j] = x_merge_path(p, p.seg, p.seg_len, p.own, p.own_len, o);
}
NLZM_KERNEL_1D(dc_xpartition, DcParams)

// Same tile scheme as dc_merge_tile_cta, query only: the own elements look up their two neighbours in the
// segment and walk the segment's chains; nothing is stored except the best lengths they reach.
DEV void dc_xmerge_tile_cta(const DcParams &p, u32 bid, u32 tid, u8 *smem) {
    Elem *el = (Elem *)smem;
    u16 *src = (u16 *)(smem + (NLZM_MT_TILE + 2) * 32);
    u16 *lcnt = src + NLZM_MT_TILE;
    u16 *act = lcnt + 2 * NLZM_MT_TILE;
    u32 *act_n = (u32 *)(act + NLZM_MT_TILE);

    const u32 total = p.seg_len + p.own_len;
    const u32 o0 = bid * NLZM_MT_TILE;
    const u32 o1 = (o0 + NLZM_MT_TILE) < total ? (o0 + NLZM_MT_TILE) : total;
    const u32 cnt = o1 - o0;
    const u32 l0 = p.part[bid];
    const u32 l1 = (o1 < total) ? p.part[bid + 1] : p.seg_len;
    const u32 r0 = o0 - l0, r1 = o1 - l1;
    const u32 nl = l1 - l0, nr = r1 - r0;
    if (nr == 0) return;                                        // no own element in this tile: nothing to ask
    Elem *SL = el + 1;
    Elem *SR = el + nl + 2;
    {
        const V16 *GL = (const V16 *)p.seg, *GR = (const V16 *)p.own;
        V16 *S = (V16 *)el;
        for (u32 c = tid; c < (nl + 2) * 2; c += NLZM_MT_THREADS) {
            const i64 gi = (i64)l0 + (i64)(c >> 1) - 1;
            if (gi >= 0 && gi < (i64)p.seg_len) S[c] = GL[gi * 2 + (c & 1)];
        }
        V16 *S2 = (V16 *)SR;
        for (u32 c = tid; c < nr * 2; c += NLZM_MT_THREADS) S2[c] = GR[(u64)r0 * 2 + c];
    }
    if (tid == 0) *act_n = 0;
    NLZM_CTA_SYNC();
    {
        const u32 t_d = tid * NLZM_MT_ITEMS < cnt ? tid * NLZM_MT_ITEMS : cnt;
        const u32 t_e = t_d + NLZM_MT_ITEMS < cnt ? t_d + NLZM_MT_ITEMS : cnt;
        u32 li = t_d < t_e ? x_merge_path(p, SL, nl, SR, nr, t_d) : 0u, ri = t_d - li;
        for (u32 o = t_d; o < t_e; o++) {
            const bool take_l = (ri >= nr) || (li < nl && x_left_first(p, SL[li], SR[ri]));
            if (take_l) {
                src[o] = (u16)(1 + li);
                ++li;
            } else {
                src[o] = (u16)(nl + 2 + ri);
                lcnt[o] = (u16)li;
                ++ri;
            }
        }
    }
    NLZM_CTA_SYNC();
    const u32 delta = (u32)(p.u0 - p.seg_u0);
    for (u32 o = tid; o < cnt; o += NLZM_MT_THREADS) {
        const u32 slot = src[o];
        if (slot < nl + 2) continue;
        const Elem &e = el[slot];
        const u64 a_abs = p.u0 + (u32)e.key;
        u32 cap;
        const u32 best_in = elem_best(e.tail);
        if (!dc_query_cap(p, a_abs, cap) || best_in >= cap || a_abs - p.seg_last > p.g.W - 1) continue;
        const u32 li = lcnt[o];
        bool want = false;
        if (l0 + li > 0) { u32 l = elem_prefix_lcp(e, el[li]); l = l < cap ? l : cap; want |= (l > best_in) || (l >= NLZM_ELEM_PREFIX && cap > NLZM_ELEM_PREFIX); }
        if (l0 + li < p.seg_len) { u32 l = elem_prefix_lcp(e, el[li + 1]); l = l < cap ? l : cap; want |= (l > best_in) || (l >= NLZM_ELEM_PREFIX && cap > NLZM_ELEM_PREFIX); }
        if (want) act[nlzm_atomic_add(act_n, 1u)] = (u16)o;
    }
    NLZM_CTA_SYNC();
    const u32 n_act = *act_n;
    for (u32 k = tid; k < n_act; k += NLZM_MT_THREADS) {
        const u32 o = act[k];
        const u32 slot = src[o];
        const Elem ea = el[slot];
        const u64 a_abs = p.u0 + (u32)ea.key;
        u32 cap = 0;
        dc_query_cap(p, a_abs, cap);
        const u32 best_in = elem_best(ea.tail);
        const u32 li = lcnt[o];
        u32 nb = best_in;
        const Elem *fl = (l0 + li > 0) ? &el[li] : nullptr, *fr = (l0 + li < p.seg_len) ? &el[li + 1] : nullptr;
        const u32 ll = fl ? elem_pair_lcp2(p, ea, p.u0, *fl, p.seg_u0, cap) : 0u, lr = fr ? elem_pair_lcp2(p, ea, p.u0, *fr, p.seg_u0, cap) : 0u;
        const u32 cl = fl ? (u32)fl->key : 0u, cr = fr ? (u32)fr->key : 0u;
        dc_walk(p, ea, a_abs, best_in, fl, ll, true, cr, lr > best_in ? lr : 0u, nb, p.seg_ptr, delta);
        dc_walk(p, ea, a_abs, best_in, fr, lr, false, cl, ll > best_in ? ll : 0u, nb, p.seg_ptr, delta);
        if (nb != best_in) p.own[r0 + (slot - (nl + 2))].tail = elem_set_best(ea.tail, nb);
    }
}
NLZM_KERNEL_CTA_OCC(dc_xmerge_tile, DcParams, NLZM_MT_THREADS, 3)

// ---- keeping only the later positions of a segment (nlzm_mf_trim_segments): order-preserving compaction
struct SegTrimParams { const Elem *in; Elem *out; u32 *flag; const u32 *idx; u32 cut; };   // cut: first kept position (relative)
DEV void seg_trim_flag_body(const SegTrimParams &p, u64 i) { p.flag[i] = (u32)p.in[i].key >= p.cut ? 1u : 0u; }
NLZM_KERNEL_1D(seg_trim_flag, SegTrimParams)
DEV void seg_trim_move_body(const SegTrimParams &p, u64 i) { if (p.flag[i]) p.out[p.idx[i]] = p.in[i]; }
NLZM_KERNEL_1D(seg_trim_move, SegTrimParams)
