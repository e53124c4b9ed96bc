// Stage M — candidate tuples -> per-position staircases (the hand-off records).
//
// Every finder appends (position, distance, length) candidates that are true matches. Repeated
// MatchTable::Update calls (NLZM.cpp:835-852) are a pointwise min-merge, so the table a position
// ends with is the lower envelope of its candidates: a candidate survives iff no other candidate
// of the same position has distance <= its distance and length >= its length. Tuples are radix
// sorted by (position, length); survivors are compacted into CSR form.
#pragma once
#include "common.cuh"

struct Step { u16 dist_lo, dist_hi, len; };        // mirrors nlzm_mf_step in include/nlzm_mf.h (6 bytes)

struct FilterParams {
    const u64 *keys;      // sorted: (a_rel << 9) | len
    const u32 *dist;
    u32 n;
    u32 *keep;            // 0/1 per tuple
    u32 *count;           // survivors per position (n_own + 1 entries, zeroed)
};
DEV void step_filter_body(const FilterParams &p, u64 j64) {
    const u32 j = (u32)j64;
    const u64 k = p.keys[j];
    const u64 a = k >> 9;
    const u32 d = p.dist[j];
    bool dead = false;
    // later tuples of the same position have length >= mine: any distance <= mine kills me
    for (u32 i = j + 1; i < p.n && (p.keys[i] >> 9) == a; i++)
        if (p.dist[i] <= d) { dead = true; break; }
    // earlier tuples with the SAME length and a strictly smaller distance also dominate
    if (!dead)
        for (u32 i = j; i-- > 0 && p.keys[i] == k; )
            if (p.dist[i] < d) { dead = true; break; }
    p.keep[j] = dead ? 0u : 1u;
    if (!dead) nlzm_atomic_add(p.count + a, 1u);
}
NLZM_KERNEL_1D(step_filter, FilterParams)

struct CompactParams { const u64 *keys; const u32 *dist; const u32 *keep; const u32 *out_idx; Step *steps; };
DEV void step_compact_body(const CompactParams &p, u64 j) {
    if (!p.keep[j]) return;
    const u32 d = p.dist[j];
    Step s; s.dist_lo = (u16)d; s.dist_hi = (u16)(d >> 16); s.len = (u16)(p.keys[j] & 511u);
    p.steps[p.out_idx[j]] = s;
}
NLZM_KERNEL_1D(step_compact, CompactParams)
