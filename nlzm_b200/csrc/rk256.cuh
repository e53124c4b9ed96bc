// Stage R — RK256 long-range finder (NLZM.cpp:1033-1123) in order-independent form.
//
//  1. hash of every 256-aligned block  H(s) = sum x[s+i] * ADDH^(256-i)  (NLZM.cpp:798-799)
//  2. "table build": blocks radix-sorted by slot (stable => time order) + CSR offsets; the one-entry
//     table of the reference holds, at time a, the last block of the slot that starts before a
//  3. per position: rolling hash, slot lookup, check bits / window test exactly as written
//     (entries keep the full shifted position, so bits >= hist_bits leak into the check field;
//     an empty slot is the raw word 0xFFFFFFFF and is matched like any entry) -> raw hit list
//  4. extension of every raw hit to its full length (cap = uint16(remaining), NLZM.cpp:759-760,1096)
//  5. the carried-match state machine over the sparse, position-sorted hit list (NLZM.cpp:1056-1069,
//     1090-1107): which hits are looked up at all, which are accepted, how long each carry lives
//  6. expansion of the carry intervals into per-position candidates
#pragma once
#include "common.cuh"
#include "dc_levels.cuh"

HD u32 rk_hash_block(const u8 *__restrict__ x, u64 s) {
    u32 h = 0;
    for (u32 i = 0; i < NLZM_RK_BLOCK; i += 8) {
        u64 w = load8(x, s + i);
        #pragma unroll
        for (int b = 0; b < 8; b++) h = ((u32)((w >> (8 * b)) & 0xFF) + h) * NLZM_RK_ADDH;   // rolling_hash_add
    }
    return h;
}

struct RkBlockParams { const u8 *x; u32 *hblk; u32 *slot_keys; u32 *vals; u32 *slot_count; u32 shift; };
DEV void rk_block_body(const RkBlockParams &p, u64 j) {
    u32 h = rk_hash_block(p.x, j * NLZM_RK_BLOCK);
    p.hblk[j] = h;
    p.slot_keys[j] = h >> p.shift;
    p.vals[j] = (u32)j;
    nlzm_atomic_add(p.slot_count + (h >> p.shift), 1u);
}
NLZM_KERNEL_1D(rk_block, RkBlockParams)

struct RkLookupParams {
    const u8 *x;
    Geom g;
    const u32 *hblk;        // hash per aligned block
    const u32 *slot_off;    // CSR offsets per slot (2^rk_bits + 1)
    const u32 *slot_blk;    // block indices sorted by (slot, block)
    u64 rk_b;               // first position looked up
    u64 rk_e;               // one past the last position looked up
    u32 span;               // positions per thread
    u64 *hit_keys;          // absolute position of the raw hit
    u32 *hit_vals;          // its distance
    u32 *hit_count;
    u32 hit_cap;
};

// One thread rolls the hash over `span` consecutive positions (NLZM.cpp:799) and does the lookups.
DEV void rk_lookup_body(const RkLookupParams &p, u64 t) {
    u64 a = p.rk_b + t * p.span;
    if (a >= p.rk_e) return;
    u64 a_end = a + p.span < p.rk_e ? a + p.span : p.rk_e;
    const u32 shift = 32 - p.g.rk_bits;
    const u32 cmask = (1u << (32 - p.g.hb)) - 1;
    u32 h = rk_hash_block(p.x, a);
    for (;; ) {
        const u32 slot = h >> shift;
        // last block of this slot that starts before a (the insert at a itself comes after the lookup)
        // (binary search: a slot can hold very many blocks, e.g. every block of a zero run)
        u32 lo = p.slot_off[slot], hi = p.slot_off[slot + 1];
        u32 e = NLZM_NONE32;
        const u32 first = lo;
        while (lo < hi) {                                        // first list entry whose block starts at or after a
            const u32 mid = lo + ((hi - lo) >> 1);
            if ((u64)p.slot_blk[mid] * NLZM_RK_BLOCK < a) lo = mid + 1; else hi = mid;
        }
        if (lo > first) {
            const u32 j = p.slot_blk[lo - 1];
            const u64 s = (u64)j * NLZM_RK_BLOCK;
            e = geom_P(p.g, s) | (p.hblk[j] << p.g.hb);          // NLZM.cpp:1111: full P, spills into the check bits
        }
        const u32 P = geom_P(p.g, a);
        const u32 sp = e & (p.g.W - 1);
        if ((e >> p.g.hb) == (h & cmask) && sp < P && P - sp <= p.g.W - 1) {     // NLZM.cpp:1091-1095
            u32 idx = nlzm_atomic_add(p.hit_count, 1u);
            if (idx < p.hit_cap) { p.hit_keys[idx] = a; p.hit_vals[idx] = P - sp; }
        }
        if (++a >= a_end) break;
        h = ((u32)p.x[a + NLZM_RK_BLOCK - 1] + h - (u32)p.x[a - 1] * NLZM_RK_REMH) * NLZM_RK_ADDH;   // rolling_hash_add_remove
    }
}
NLZM_KERNEL_1D(rk_lookup, RkLookupParams)

struct RkExtendParams {
    const u8 *x; Geom g;
    const u64 *hit_pos; const u32 *hit_dist; u32 *hit_len;
    u64 *valid_pos; u32 *valid_idx; u32 *valid_count;     // hits that pass mlen >= match_min(dist), unsorted
};
DEV void rk_extend_body(const RkExtendParams &p, u64 i) {
    const u64 a = p.hit_pos[i];
    const u32 d = p.hit_dist[i];
    const u32 cap = geom_rem(p.g, a) & 0xFFFFu;          // uint16 max_len parameter, NLZM.cpp:759-760,1096-1097
    const u32 m = lcp_cap(p.x, a - d, a, cap);
    p.hit_len[i] = m;
    if (m >= match_min(d)) {                             // everything else is not a hit at all (NLZM.cpp:1099)
        u32 j = nlzm_atomic_add(p.valid_count, 1u);
        p.valid_pos[j] = a;
        p.valid_idx[j] = (u32)i;
    }
}
NLZM_KERNEL_1D(rk_extend, RkExtendParams)

#ifndef NLZM_EMU
// Warp-cooperative version used on the GPU: one warp per raw hit. Each lane compares 8 bytes of a
// 256-byte stripe per step (coalesced on both streams), the first mismatching lane is found with
// __ballot_sync / __ffs, and the byte inside it with a trailing-zero count. Matches run up to 65535
// bytes (NLZM.cpp:1096-1097), so a single thread per hit serialises badly on redundant data.
__global__ void __launch_bounds__(256) k_rk_extend_warp(const RkExtendParams p, u32 n_hits) {
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (warp >= n_hits) return;
    const u64 a = p.hit_pos[warp];
    const u32 d = p.hit_dist[warp];
    const u32 cap = geom_rem(p.g, a) & 0xFFFFu;
    u32 m = 0;
    while (m < cap) {
        const u32 off = m + lane * 8;
        u64 diff = 0;
        if (off < cap) diff = load8(p.x, a - d + off) ^ load8(p.x, a + off);      // bytes past cap are ignored below
        const unsigned bad = __ballot_sync(0xFFFFFFFFu, diff != 0);
        if (bad) {
            const int first = __ffs(bad) - 1;
            const u64 dd = __shfl_sync(0xFFFFFFFFu, diff, first);
            m += (u32)first * 8 + ((u32)(__ffsll((long long)dd) - 1) >> 3);
            break;
        }
        m += 256;
    }
    if (m > cap) m = cap;
    if (lane == 0) {
        p.hit_len[warp] = m;
        if (m >= match_min(d)) {
            u32 j = atomicAdd(p.valid_count, 1u);
            p.valid_pos[j] = a;
            p.valid_idx[j] = warp;
        }
    }
}
static inline void launch_rk_extend_warp(const RkExtendParams &p, u64 n_hits, cudaStream_t st) {
    if (n_hits == 0) return;
    nlzm_launch_begin("k_rk_extend_warp", st);
    k_rk_extend_warp<<<(unsigned)((n_hits * 32 + 255) / 256), 256, 0, st>>>(p, (u32)n_hits);
    nlzm_launch_end(st);
}
#endif

struct RkInterval { u64 start; u32 dist; u32 len; u64 end; };

struct RkChainParams {
    Geom g;
    const u64 *valid_pos; const u32 *valid_idx;          // valid hits sorted by position
    const u32 *hit_dist; const u32 *hit_len;             // indexed by raw hit index
    const u32 *n_valid;
    RkInterval *iv; u32 *n_iv;
    // Where the machine may start. A carry lives at most 65535 positions (uint16 length, NLZM.cpp:759-760), so after
    // 65536 positions without a valid hit nothing is carried: the state is "no carry" whatever came before. When the
    // hits were only looked up from `known_from` on (not from the last ring shift), the walk starts behind the latest
    // such stretch that ends at or before `own_b`; if there is none, *ok = 0 and the caller redoes the range in full.
    u64 known_from, own_b;
    u32 need_restart;
    u32 *ok;
};
#define NLZM_RK_CARRY_MAX 65536ull

HD u64 rk_carry_end(const Geom &g, u64 ca, u32 cl) {
    // a carry dies when its length is used up or at the next ring shift (P - carry_to wraps, NLZM.cpp:1057)
    u64 end = ca + cl;
    const u32 ep = geom_epoch(g, ca);
    // first chunk start after ca whose epoch differs: epochs only change at chunk starts
    u64 k = ca / g.cs + 1;
    // a carry spans at most 65535 positions => at most a few chunks
    while (k * g.cs < end) {
        if (geom_epoch(g, k * g.cs) != ep) { end = k * g.cs; break; }
        ++k;
    }
    return end;
}

// The sequential part: one thread walks the position-sorted hits.
DEV void rk_chain_body(const RkChainParams &p, u64) {
    const u32 n = *p.n_valid;
    u32 cl = 0, cd = 0, cep = 0, niv = 0;
    u64 ca = 0;
    u32 i = 0;
    *p.ok = 1;
    if (p.need_restart) {
        u64 last = p.known_from;                         // nothing is known about hits before this position
        u32 start = 0xFFFFFFFFu, j = 0;
        for (; j < n && p.valid_pos[j] < p.own_b; j++) {
            if (p.valid_pos[j] - last >= NLZM_RK_CARRY_MAX) start = j;
            last = p.valid_pos[j];
        }
        if (p.own_b - last >= NLZM_RK_CARRY_MAX) start = j;
        if (start == 0xFFFFFFFFu) { *p.ok = 0; *p.n_iv = 0; return; }
        i = start;
    }
    while (i < n) {
        const u64 a = p.valid_pos[i];
        const u32 ri = p.valid_idx[i];
        const u32 d = p.hit_dist[ri], m = p.hit_len[ri];
        const bool alive = cl > 0 && geom_epoch(p.g, a) == cep && a - ca < cl;
        if (alive && cl >= NLZM_RK_BLOCK) {
            // no lookups under a long carry (NLZM.cpp:1090): skip straight to the first hit at or after its end
            const u64 end = rk_carry_end(p.g, ca, cl);
            u32 lo = i + 1, hi = n;
            while (lo < hi) { const u32 mid = lo + ((hi - lo) >> 1); if (p.valid_pos[mid] < end) lo = mid + 1; else hi = mid; }
            i = lo;
            continue;
        }
        if (alive && m < cl) { ++i; continue; }                           // must be >= the carry's original length (1099)
        if (cl > 0) {
            u64 end = rk_carry_end(p.g, ca, cl);
            RkInterval v; v.start = ca; v.dist = cd; v.len = cl; v.end = (a + 1 < end) ? a + 1 : end;   // (A) still fires at a
            p.iv[niv++] = v;
        }
        ca = a; cd = d; cl = m; cep = geom_epoch(p.g, a);
        ++i;
    }
    if (cl > 0) {
        RkInterval v; v.start = ca; v.dist = cd; v.len = cl; v.end = rk_carry_end(p.g, ca, cl);
        p.iv[niv++] = v;
    }
    *p.n_iv = niv;
}
NLZM_KERNEL_1D(rk_chain, RkChainParams)

struct RkExpandParams { Geom g; const RkInterval *iv; u64 own_b, own_e; u32 bt_on; TupleSink sink; };
DEV void rk_expand_body(const RkExpandParams &p, u64 i) {
    const RkInterval v = p.iv[i];
    const u32 mm = match_min(v.dist);
    u64 a = v.start > p.own_b ? v.start : p.own_b;
    u64 end = v.end < p.own_e ? v.end : p.own_e;
    if (p.bt_on) {
        // exhaustive BT4 reports a candidate at least as near and as long for every match of 4+ bytes
        // (see HtFindParams): only the last positions of the carry, where 2..3 bytes remain, can add something
        const u64 tail = v.start + (v.len > 3 ? v.len - 3 : 0);
        if (tail > a) a = tail;
    }
    for (; a < end; a++) {
        if (p.g.flen - a < NLZM_RK_BLOCK) break;                      // RK is not called there (NLZM.cpp:1525)
        u32 r = v.len - (u32)(a - v.start);
        if (r < mm) break;                                            // only shrinks from here on
        tuple_append(p.sink, (u32)(a - p.own_b), v.dist, r < NLZM_MATCH_MAX ? r : NLZM_MATCH_MAX);
    }
}
NLZM_KERNEL_1D(rk_expand, RkExpandParams)
