// Stage S — suffix ranks of the universe [u0, u0+n) to depth >= 264 bytes by prefix doubling.
// Output: rank[i] (1-based index of the group head in sorted order) such that comparing
// (rank[i], i) as a u64 is a strict total order consistent with comparing the first 264 bytes of
// the suffixes (zero padded past the end of the file). The finders only need this ORDER: a set
// {q : lcp(q, a) >= L} is then a contiguous interval around a.
#pragma once
#include "common.cuh"
#include "prim.cuh"

struct RankInitParams { const u8 *x; u64 u0; u64 *keys; u32 *vals; };
DEV void rank_init_body(const RankInitParams &p, u64 i) {
    p.keys[i] = bswap64(load8(p.x, p.u0 + i));     // first 8 bytes, big endian => numeric order == byte order
    p.vals[i] = (u32)i;
}
NLZM_KERNEL_1D(rank_init, RankInitParams)

struct RankHeadParams { const u64 *keys; u32 *head_idx; u32 *head_flag; };
DEV void rank_head_body(const RankHeadParams &p, u64 j) {
    bool head = (j == 0) || (p.keys[j] != p.keys[j - 1]);
    p.head_idx[j] = head ? (u32)j : 0u;
    p.head_flag[j] = head ? 1u : 0u;
}
NLZM_KERNEL_1D(rank_head, RankHeadParams)

struct RankScatterParams { const u32 *vals; const u32 *group_head; u32 *rank; };
DEV void rank_scatter_body(const RankScatterParams &p, u64 j) { p.rank[p.vals[j]] = p.group_head[j] + 1u; }
NLZM_KERNEL_1D(rank_scatter, RankScatterParams)

struct RankPairParams { const u32 *rank; u64 *keys; u32 *vals; u64 n; u64 h; };
DEV void rank_pair_body(const RankPairParams &p, u64 i) {
    u32 r2 = (i + p.h < p.n) ? p.rank[i + p.h] : 0u;   // past the universe: sorts first
    p.keys[i] = ((u64)p.rank[i] << 32) | r2;
    p.vals[i] = (u32)i;
}
NLZM_KERNEL_1D(rank_pair, RankPairParams)

