// BT4 candidates of length 2..3 (NLZM.cpp:996: a tree node is reported whenever
// mlen >= get_match_min(distance), also when it only shares the 4-byte-hash BUCKET with the
// current position). Two positions that agree on 2 or 3 bytes but not on 4 can share a bucket
// only if the bucket index has fewer than 16 bits (hist_bits < 19), and the
// candidate matters only while it is nearer than the nearest >= 4-byte match, so the walk over
// the bucket's earlier positions stops at the first of those or 4095 bytes back.
#pragma once
#include "common.cuh"
#include "dc_levels.cuh"

struct BtBucketParams { const u8 *x; u64 s0; u32 shift; u32 *keys; u32 *vals; };
DEV void bt_bucket_body(const BtBucketParams &p, u64 i) {
    p.keys[i] = (load4(p.x, p.s0 + i) * NLZM_HASH_MUL) >> p.shift;
    p.vals[i] = (u32)i;
}
NLZM_KERNEL_1D(bt_bucket, BtBucketParams)

struct BtShortParams {
    const u8 *x; Geom g;
    const u32 *skeys; const u32 *svals; const u32 *inv;   // positions [s0, ..) sorted by bucket
    u64 s0; u64 own_b;
    TupleSink sink;
};
DEV void bt_short_body(const BtShortParams &p, u64 i) {
    const u64 a = p.own_b + i;
    const u32 me = (u32)(a - p.s0);
    u32 j = p.inv[me];
    const u32 bucket = p.skeys[j];
    const u32 v4 = load4(p.x, a);
    u32 seen = 0;
    while (j > 0 && p.skeys[j - 1] == bucket) {
        --j;
        const u64 q = p.s0 + p.svals[j];
        const u32 d = (u32)(a - q);
        if (d > 4095 || d > p.g.W - 1) break;
        const u32 diff = load4(p.x, q) ^ v4;
        if (diff == 0) break;                              // a >= 4-byte match: everything farther is dominated
        const u32 l = (u32)nlzm_ctz64((u64)diff) >> 3;
        if (l <= seen || l < match_min(d)) continue;
        tuple_append(p.sink, (u32)i, d, l);
        seen = l;
        if (seen == 3) break;
    }
}
NLZM_KERNEL_1D(bt_short, BtShortParams)
