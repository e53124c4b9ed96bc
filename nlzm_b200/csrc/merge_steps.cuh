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

// Pre-pricing (NLZM.cpp:1556-1596): what the parser derives from a candidate's distance before it can
// price it — the distance slot of the stream's distance code (NLZM.cpp:1219-1236; the raw-bit count follows from the
// slot) and the shortest length a match at this distance may have (get_match_min, NLZM.cpp:813-821) — is computed
// here, once per emitted step, and travels in the spare bits of the 6-byte record.
HD u32 step_dist_slot(u32 dist) {
    const u32 v = dist - 1;
    if (v < 4) return v;
    const u32 nb = 32u - (u32)
#ifdef __CUDA_ARCH__
        __clz((int)v);
#else
        __builtin_clz(v);
#endif
    return ((nb - 1) << 1) + ((v >> (nb - 2)) & 1u);
}

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
    Step s;
    s.dist_lo = (u16)d;
    s.dist_hi = (u16)((d >> 16) | ((match_min(d) - 2u) << 12));            // bits 12..13: shortest length - 2
    s.len = (u16)((p.keys[j] & 511u) | (step_dist_slot(d) << 9));          // bits 9..14: distance slot
    p.steps[p.out_idx[j]] = s;
}
NLZM_KERNEL_1D(step_compact, CompactParams)
