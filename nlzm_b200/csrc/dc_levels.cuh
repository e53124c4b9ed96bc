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
#pragma once
#include "common.cuh"

struct PtrEntry {
    u64 pg;   // key (rank<<32|pos) of the nearest element to the rank-left with a greater position in
              // the same segment, 0 if none
    u64 ng;   // same to the rank-right, ~0 if none
};
#define NLZM_PG_NONE 0ull
#define NLZM_NG_NONE 0xFFFFFFFFFFFFFFFFull

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

struct LevelParams {
    const u64 *cur;      // level k-1 arrays: each aligned segment of h elements sorted by key
    u64 *nxt;            // level k arrays
    u32 *corank;         // per cur index of a left-half element: its co-rank in the right half
    PtrEntry *ptr;       // indexed by universe-relative position
    u16 *best;           // indexed by universe-relative position
    const u8 *x;         // whole input, absolute
    Geom g;
    u64 u0;              // absolute offset of the universe
    u64 own_b, own_e;    // absolute range whose candidates are wanted
    u32 n;               // universe size
    u32 h;               // half segment size at this level
    TupleSink sink;
};

HD u32 lower_bound_u64(const u64 *__restrict__ a, u32 n, u64 key) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Walk one greater-position chain of the left half and emit the candidates that beat best_in.
DEV void dc_walk(const LevelParams &p, u32 a_rel_u, u64 a_abs, u32 cap, u32 best_in, u64 start_key, bool left,
                 u32 &new_best) {
    const u32 none_lo = left ? 0u : 0xFFFFFFFFu;
    if (start_key == (left ? NLZM_PG_NONE : NLZM_NG_NONE)) return;
    (void)none_lo;
    u32 c = (u32)start_key;
    u32 pend_len = 0, pend_c = 0;
    u32 lim = cap;
    const u32 a_rel = (u32)(a_abs - p.own_b);
    while (true) {
        u32 l = lcp_cap(p.x, p.u0 + c, a_abs, lim);
        if (l <= best_in) break;
        if (pend_len && l < pend_len) {
            u32 d = a_rel_u - pend_c;
            if (d <= p.g.W - 1 && pend_len >= match_min(d)) tuple_append(p.sink, a_rel, d, pend_len);
        }
        pend_len = l; pend_c = c;       // equal lcp: the nearer element replaces the farther one
        lim = l;                        // lcp never grows along the chain
        if (l > new_best) new_best = l;
        u64 k = left ? p.ptr[c].pg : p.ptr[c].ng;
        if (k == (left ? NLZM_PG_NONE : NLZM_NG_NONE)) break;
        c = (u32)k;
    }
    if (pend_len) {
        u32 d = a_rel_u - pend_c;
        if (d <= p.g.W - 1 && pend_len >= match_min(d)) tuple_append(p.sink, a_rel, d, pend_len);
    }
}

// Kernel A: merge by co-rank (binary search in the sibling half) + queries of right-half elements.
DEV void dc_merge_query_body(const LevelParams &p, u64 idx64) {
    const u32 idx = (u32)idx64;
    const u64 key = p.cur[idx];
    const u32 two_h = p.h << 1;                 // h <= 2^31 is guaranteed by the caller
    const u32 base = (idx / two_h) * two_h;
    const u32 l_len = (p.n - base) < p.h ? (p.n - base) : p.h;
    const u32 r_beg = base + l_len;
    const u32 r_len = (p.n - r_beg) < p.h ? (p.n - r_beg) : p.h;
    if (r_len == 0) { p.nxt[idx] = key; return; }
    if (idx < r_beg) {
        u32 u = lower_bound_u64(p.cur + r_beg, r_len, key);
        p.corank[idx] = u;
        p.nxt[idx + u] = key;
        return;
    }
    const u32 t = lower_bound_u64(p.cur + base, l_len, key);
    p.nxt[base + (idx - r_beg) + t] = key;

    const u32 pos = (u32)key;
    const u64 a_abs = p.u0 + pos;
    if (a_abs < p.own_b || a_abs >= p.own_e) return;
    const u64 left_in_file = p.g.flen - a_abs;
    if (left_in_file < 4) return;                                   // HT/BT need 4 visible bytes (NLZM.cpp:1515)
    const u32 cap = left_in_file < NLZM_MATCH_MAX ? (u32)left_in_file : NLZM_MATCH_MAX;   // NLZM.cpp:987
    const u32 best_in = p.best[pos];
    if (best_in >= cap) return;                                     // already matched to the cap at a nearer level
    if (pos - (r_beg - 1) > p.g.W - 1) return;                      // the whole left half is outside the window
    u32 nb = best_in;
    dc_walk(p, pos, a_abs, cap, best_in, t > 0 ? p.cur[base + t - 1] : NLZM_PG_NONE, true, nb);
    dc_walk(p, pos, a_abs, cap, best_in, t < l_len ? p.cur[base + t] : NLZM_NG_NONE, false, nb);
    if (nb != best_in) p.best[pos] = (u16)nb;
}
NLZM_KERNEL_1D(dc_merge_query, LevelParams)

// Kernel B: left-half elements adopt their rank-nearest right-half neighbours as pg / ng when those
// are nearer in rank than the pointers they already hold (every right-half position is greater).
DEV void dc_link_body(const LevelParams &p, u64 idx64) {
    const u32 idx = (u32)idx64;
    const u32 two_h = p.h << 1;
    const u32 base = (idx / two_h) * two_h;
    const u32 l_len = (p.n - base) < p.h ? (p.n - base) : p.h;
    const u32 r_beg = base + l_len;
    if (idx >= r_beg) return;
    const u32 r_len = (p.n - r_beg) < p.h ? (p.n - r_beg) : p.h;
    if (r_len == 0) return;
    const u32 pos = (u32)p.cur[idx];
    const u32 u = p.corank[idx];
    PtrEntry e = p.ptr[pos];
    bool ch = false;
    if (u > 0) { u64 k = p.cur[r_beg + u - 1]; if (k > e.pg) { e.pg = k; ch = true; } }
    if (u < r_len) { u64 k = p.cur[r_beg + u]; if (k < e.ng) { e.ng = k; ch = true; } }
    if (ch) p.ptr[pos] = e;
}
NLZM_KERNEL_1D(dc_link, LevelParams)

struct DcInitParams { PtrEntry *ptr; u16 *best; u16 best0; };
DEV void dc_init_body(const DcInitParams &p, u64 i) {
    PtrEntry e; e.pg = NLZM_PG_NONE; e.ng = NLZM_NG_NONE;
    p.ptr[i] = e;
    p.best[i] = p.best0;
}
NLZM_KERNEL_1D(dc_init, DcInitParams)
