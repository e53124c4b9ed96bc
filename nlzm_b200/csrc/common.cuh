// Shared device/host helpers: encoder geometry, minimum match length, unaligned loads, lcp.
// Reference semantics cited as NLZM.cpp:line (nauful/NLZM 1.03).
#pragma once
#include "platform.cuh"

#define NLZM_NONE32 0xFFFFFFFFu
#define NLZM_MATCH_MAX 264u                 // NLZM.cpp:737
#define NLZM_HASH_MUL 987660757u            // NLZM.cpp:739
#define NLZM_RK_BLOCK 256u                  // NLZM.cpp:789-791
#define NLZM_RK_ADDH 0x2F0FD693u            // NLZM.cpp:793
#define NLZM_RK_REMH 0x0E4EA401u            // NLZM.cpp:796 (= ADDH^256 mod 2^32)
#define NLZM_X_PAD 64u                      // readable zero bytes after the input in HBM

// Geometry of encode_file that is part of the matcher semantics (NLZM.cpp:1716-1725, 1750-1753,
// 1782-1798). All positions are absolute file offsets; the reference's shifted coordinate is
// P(a) = a - W * epoch(a).
struct Geom {
    u64 flen;
    u32 hb;        // hist_bits after the shrink loop
    u32 W;         // 1 << hb
    u32 cs;        // chunk_size
    u32 ht3_bits, bt_bits, rk_bits;
};

static inline u32 nlzm_clampu(u32 v, u32 lo, u32 hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline Geom make_geom(u64 flen, u32 hist_bits_req) {
    u32 hb = nlzm_clampu(hist_bits_req, 15, 28);                      // NLZM.cpp:2085
    while (hb > 10 && flen < (1ull << (hb - 1))) --hb;                // NLZM.cpp:1716-1718
    Geom g;
    g.flen = flen;
    g.hb = hb;
    g.W = 1u << hb;
    u32 frame_bits = nlzm_clampu(hb - 2, 14, 17);                     // NLZM.cpp:1722
    g.cs = ((1u << frame_bits) * 15) / 16 - 0x200;                    // NLZM.cpp:1724
    g.ht3_bits = 12 + nlzm_clampu(hb, 15, 17) - 15;                   // NLZM.cpp:1751
    g.bt_bits = 13 + nlzm_clampu(hb, 16, 20) - 16;                    // NLZM.cpp:1752
    g.rk_bits = 15 + nlzm_clampu(hb, 16, 22) - 16;                    // NLZM.cpp:1753
    return g;
}

// distance-dependent minimum match length, NLZM.cpp:813-821
HD u32 match_min(u32 d) { return 2u + (d >= 256u) + (d >= 4096u) + (d >= (1u << 20)); }

// number of ring shifts applied when the chunk holding a is processed (NLZM.cpp:1786-1792):
// chunk k starts at k*cs with hist_pos = k*cs - W*n_k and n_k = max(0, floor(k*cs/W) - 1).
HD u32 geom_epoch(const Geom &g, u64 a) {
    u64 k = a / g.cs;
    u64 q = (k * g.cs) >> g.hb;
    return q > 1 ? (u32)(q - 1) : 0u;
}
HD u32 geom_P(const Geom &g, u64 a) { return (u32)(a - ((u64)geom_epoch(g, a) << g.hb)); }
// bytes visible from a: the chunk's lookahead ends 265 bytes past the chunk (NLZM.cpp:1725,1797-1798)
HD u32 geom_rem(const Geom &g, u64 a) {
    u64 k = a / g.cs;
    u64 end = (k + 1) * g.cs + (NLZM_MATCH_MAX + 1);
    if (end > g.flen) end = g.flen;
    return (u32)(end - a);
}

// 8 bytes at an arbitrary offset, little endian, from an 8-byte aligned base with padding
HD u64 load8(const u8 *__restrict__ x, u64 p) {
    const u64 *w = (const u64 *)(x + (p & ~7ull));
    u32 sh = (u32)(p & 7) * 8;
    u64 lo = w[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (w[1] << (64 - sh));
}
HD u32 load4(const u8 *__restrict__ x, u64 p) { return (u32)load8(x, p); }

// common prefix length of x[p0..] and x[p1..], capped
HD u32 lcp_cap(const u8 *__restrict__ x, u64 p0, u64 p1, u32 cap) {
    u32 m = 0;
    while (m < cap) {
        u64 d = load8(x, p0 + m) ^ load8(x, p1 + m);
        if (d) {
            m += (u32)nlzm_ctz64(d) >> 3;
            break;
        }
        m += 8;
    }
    return m < cap ? m : cap;
}

HD u64 bswap64(u64 v) {
    v = ((v & 0x00FF00FF00FF00FFull) << 8) | ((v >> 8) & 0x00FF00FF00FF00FFull);
    v = ((v & 0x0000FFFF0000FFFFull) << 16) | ((v >> 16) & 0x0000FFFF0000FFFFull);
    return (v << 32) | (v >> 32);
}
