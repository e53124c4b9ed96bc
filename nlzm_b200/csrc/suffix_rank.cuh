// Stage S — suffix ranks of the universe [u0, u0+n) to depth >= 264 bytes by prefix doubling.
// Output: rank[i] (1-based index of the group head in sorted order) such that comparing
// (rank[i], i) as a u64 is a strict total order consistent with comparing the first 264 bytes of
// the suffixes (zero padded past the end of the file). The finders only need this ORDER: a set
// {q : lcp(q, a) >= L} is then a contiguous interval around a.
//
// Round 0 sorts all positions by their first 8 bytes. Every later round doubles the depth but only
// touches the ACTIVE positions — those whose group (equal prefix so far) still has two or more
// members: they are re-sorted inside their group by the rank of the position h bytes further on.
#pragma once
#include "common.cuh"
#include "prim.cuh"

struct RankInitParams { const u8 *x; u64 u0; u64 *keys; u32 *vals; };
DEV void rank_init_body(const RankInitParams &p, u64 i) {
    p.keys[i] = bswap64(load8(p.x, p.u0 + i));     // first 8 bytes, big endian => numeric order == byte order
    p.vals[i] = (u32)i;
}
NLZM_KERNEL_1D(rank_init, RankInitParams)

// keys of a refinement round: (group << nb) | rank of the suffix h bytes later (0 past the universe)
struct RankKeysParams { const u32 *grp; const u32 *pos; const u32 *rank; u64 *keys; u32 *vals; u64 n; u64 h; u32 nb; };
DEV void rank_keys_body(const RankKeysParams &p, u64 i) {
    const u32 q = p.pos[i];
    const u32 r2 = (q + p.h < p.n) ? p.rank[q + p.h] : 0u;
    p.keys[i] = ((u64)p.grp[i] << p.nb) | r2;
    p.vals[i] = q;
}
NLZM_KERNEL_1D(rank_keys, RankKeysParams)

// after the sort: start index of the old group and of the new (refined) group of every element
struct RankHeadParams { const u64 *keys; u32 *g_start; u32 *s_start; u32 nb; u32 round0; };
DEV void rank_head_body(const RankHeadParams &p, u64 j) {
    const u64 k = p.keys[j];
    bool gh = (j == 0), sh = (j == 0);
    if (j > 0) {
        const u64 pk = p.keys[j - 1];
        sh = pk != k;
        gh = p.round0 ? false : ((pk >> p.nb) != (k >> p.nb));
    }
    p.g_start[j] = gh ? (u32)j : 0u;     // inclusive max-scan turns these into "index of my group's first element"
    p.s_start[j] = sh ? (u32)j : 0u;
}
NLZM_KERNEL_1D(rank_head, RankHeadParams)

// new rank = old group's global start + offset of the refined group inside it; still-ambiguous
// elements (refined group of >= 2) stay active
struct RankAssignParams {
    const u64 *keys; const u32 *vals; const u32 *g_start; const u32 *s_start;
    u32 *rank; u32 *new_grp; u32 *act_flag; u64 m; u32 nb; u32 round0;
};
DEV void rank_assign_body(const RankAssignParams &p, u64 j) {
    const u32 grp = p.round0 ? 0u : (u32)(p.keys[j] >> p.nb);
    const u32 base = grp + (p.s_start[j] - p.g_start[j]);
    p.rank[p.vals[j]] = base + 1u;
    p.new_grp[j] = base;
    const bool head = p.s_start[j] == (u32)j;
    const bool next_head = (j + 1 == p.m) || (p.s_start[j + 1] == (u32)(j + 1));
    p.act_flag[j] = (head && next_head) ? 0u : 1u;
}
NLZM_KERNEL_1D(rank_assign, RankAssignParams)

struct RankCompactParams { const u32 *act_flag; const u32 *act_idx; const u32 *new_grp; const u32 *vals; u32 *grp_out; u32 *pos_out; };
DEV void rank_compact_body(const RankCompactParams &p, u64 j) {
    if (!p.act_flag[j]) return;
    const u32 o = p.act_idx[j];
    p.grp_out[o] = p.new_grp[j];
    p.pos_out[o] = p.vals[j];
}
NLZM_KERNEL_1D(rank_compact, RankCompactParams)
