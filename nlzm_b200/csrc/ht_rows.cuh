// Stage H — HT2 / HT3 short-match finders (NLZM.cpp:893-957) without the tables.
//
// The reference keeps `rows << bits` u32 cells; an access with bucket b reads cells b .. b+rows-1
// (rows overlap because the row pointer is `rows + bucket`, NLZM.cpp:912), then writes its own
// entry into cell b and pushes the old content of cell b into cell b+1. What a reader sees in a
// cell is "the last write before me", and a cell c is written by accesses of bucket c (own entry)
// and of bucket c-1 (the pushed old content of cell c-1). So everything follows from three numbers
// per position p with bucket b:  PS[p], PL[p], PR[p] = the last access before p in bucket b, b-1,
// b+1. They are computed with per-tile "last access" tables:
//   k_ht_tile_last   per tile of 32 Ki positions: last access per bucket (shared-memory atomicMax)
//   k_ht_scan_*      exclusive running max over the tiles, per bucket, in two levels (groups of 64 tiles)
//                    => table at every tile start
//   k_ht_prev        per tile, in position order: one warp walks the tile 32 positions at a time with a
//                    u16 table (in-tile offsets) and 2 KiB of staged text in shared memory; accesses
//                    inside the same 32 are resolved from the __match_any mask and the table itself,
//                    with 31 shuffles only when a later lane touched a neighbouring bucket
//   k_ht_find        per position: resolve the raw cell contents (a short chain when the last writer
//                    pushed an older entry), then the reference's match / check-bit / window logic
// MatchFinderHT::Shift as written only clears cell 0 at each ring shift (NLZM.cpp:940-957); tables
// never age, so the last-access tables cover the whole prefix [0, end) — streaming work, no sort.
// PS/PL/PR are materialised from `pos0` on: by default 0 (the whole prefix, 12 B per position). With the
// "ht_margin" option they start that far before the answered range and the prefix before that only gets
// coarse tables per 2^ht_coarse_log positions (table before the tile, first / last access and access
// count inside it); a chain step that lands there is answered exactly from those, scanning text only when
// the position lies between accesses of a bucket with three or more accesses in its tile. Measured on a
// late shard of 800 MB text with an 8 Mi margin: those scans cost 70 ms, the whole-prefix walk 27 ms.
#pragma once
#include "common.cuh"
#include "dc_levels.cuh"

struct HtCfg {
    u32 rows;      // 1 (HT2) or 2 (HT3)
    u32 bits;      // bucket index bits
    u32 nbytes;    // 2 or 3 hashed bytes
};

#define NLZM_HT_TILE_LOG 15u                   // fine tiles: k_ht_prev walks these in order (in-tile offsets fit u16)
#define NLZM_HT_GROUP 64u                      // tiles per group of the two-level running max
#define NLZM_HT_TILE (1u << NLZM_HT_TILE_LOG)
#define NLZM_HT_COARSE_LOG 17u                 // coarse tiles of the far prefix
#define NLZM_HT_MARGIN 0ull                    // PS/PL/PR start this far before the answered range (rounded down to a coarse
                                               // tile): chains that reach further back end in the snapshot of the cells
                                               // at that point, so the per-position walk does not grow with the prefix
#ifndef NLZM_HT_THREADS
#define NLZM_HT_THREADS 256
#endif
#define NLZM_HT_STAGE 2048u                     // text bytes staged in shared memory by k_ht_prev

HD u32 ht_hash(const u8 *__restrict__ x, u64 a, u32 nbytes) {
    u32 v = load4(x, a) & (nbytes == 2 ? 0xFFFFu : 0xFFFFFFu);      // VALUE2 / VALUE3, NLZM.cpp:741-742
    return v * NLZM_HASH_MUL;
}

struct HtTableParams {
    const u8 *x;
    HtCfg c;
    u64 pos0;         // first position the tiles of this launch cover (multiple of the tile size)
    u64 n_acc;        // accesses happen at positions [0, n_acc) (4 bytes visible, NLZM.cpp:1515); tiles end here
    u64 x_limit;      // 8-byte words may be read at offsets <= x_limit (input length + padding - 8)
    u32 tile_log;     // log2 of the tile size of this launch
    u32 n_tiles;
    u32 *tile_last;   // [n_tiles][1 << bits]: last access (+1) per bucket inside the tile, then: before the tile
    u32 *first_rows;  // out (coarse launch only, else null): [n_tiles][1 << bits] first access (+1) inside the tile
    u32 *last_rows;   // out (coarse launch only): copy of the per-tile last access before the scan overwrites it
    u32 *count_rows;  // out (coarse launch only): accesses per bucket inside the tile
    u32 *ps, *pl, *pr;// per position - pos0: last access (+1, 0 = none) before it in bucket b, b-1, b+1
};

// ---- per tile: last access per bucket
DEV void ht_tile_last_cta(const HtTableParams &p, u32 bid, u32 tid, u8 *smem) {
    u32 *tab = (u32 *)smem;
    const u32 nc = 1u << p.c.bits;
    u32 *tmin = tab + nc, *tcnt = tab + 2 * nc;                  // only with first_rows (launcher sizes the memory)
    for (u32 i = tid; i < nc; i += NLZM_HT_THREADS) { tab[i] = 0; if (p.first_rows) { tmin[i] = 0xFFFFFFFFu; tcnt[i] = 0; } }
    NLZM_CTA_SYNC();
    const u64 t0 = p.pos0 + ((u64)bid << p.tile_log);
    const u64 t1 = t0 + (1ull << p.tile_log) < p.n_acc ? t0 + (1ull << p.tile_log) : p.n_acc;
    for (u64 a = t0 + tid; a < t1; a += NLZM_HT_THREADS) {
        const u32 b = ht_hash(p.x, a, p.c.nbytes) >> (32 - p.c.bits);
        nlzm_atomic_max(tab + b, (u32)a + 1u);
        if (p.first_rows) { nlzm_atomic_min(tmin + b, (u32)a + 1u); nlzm_atomic_add(tcnt + b, 1u); }
    }
    NLZM_CTA_SYNC();
    u32 *out = p.tile_last + (u64)bid * nc;
    for (u32 i = tid; i < nc; i += NLZM_HT_THREADS) {
        out[i] = tab[i];
        if (p.first_rows) {
            p.last_rows[(u64)bid * nc + i] = tab[i];
            p.first_rows[(u64)bid * nc + i] = tmin[i] == 0xFFFFFFFFu ? 0u : tmin[i];
            p.count_rows[(u64)bid * nc + i] = tcnt[i];
        }
    }
}
NLZM_KERNEL_CTA(ht_tile_last, HtTableParams, NLZM_HT_THREADS)

// ---- exclusive running max over the tiles, per bucket, in two levels so that long prefixes stay parallel:
//      group maxima -> running max over groups -> running max inside each group
struct HtScanParams {
    u32 *tile_last;     // [n_tiles][nc], converted in place to "table before the tile"
    u32 *group_max;     // [n_groups][nc] scratch
    const u32 *init;    // table before the first tile, or null
    u32 *final_row;     // out: table after the last tile, or null
    u32 nc, n_tiles, n_groups;
};
DEV void ht_scan_group_body(const HtScanParams &p, u64 i) {         // i = group * nc + bucket
    const u32 g = (u32)(i / p.nc), b = (u32)(i % p.nc);
    const u32 t1 = (g + 1) * NLZM_HT_GROUP < p.n_tiles ? (g + 1) * NLZM_HT_GROUP : p.n_tiles;
    u32 m = 0;
    for (u32 t = g * NLZM_HT_GROUP; t < t1; t++) { const u32 v = p.tile_last[(u64)t * p.nc + b]; m = v > m ? v : m; }
    p.group_max[i] = m;
}
NLZM_KERNEL_1D(ht_scan_group, HtScanParams)
DEV void ht_scan_top_body(const HtScanParams &p, u64 b) {           // one thread per bucket over the groups
    u32 run = p.init ? p.init[b] : 0u;
    for (u32 g = 0; g < p.n_groups; g++) {
        u32 *cell = p.group_max + (u64)g * p.nc + b;
        const u32 v = *cell;
        *cell = run;
        run = v > run ? v : run;
    }
    if (p.final_row) p.final_row[b] = run;
}
NLZM_KERNEL_1D(ht_scan_top, HtScanParams)
DEV void ht_scan_apply_body(const HtScanParams &p, u64 i) {
    const u32 g = (u32)(i / p.nc), b = (u32)(i % p.nc);
    const u32 t1 = (g + 1) * NLZM_HT_GROUP < p.n_tiles ? (g + 1) * NLZM_HT_GROUP : p.n_tiles;
    u32 run = p.group_max[i];
    for (u32 t = g * NLZM_HT_GROUP; t < t1; t++) {
        u32 *cell = p.tile_last + (u64)t * p.nc + b;
        const u32 v = *cell;
        *cell = run;
        run = v > run ? v : run;
    }
}
NLZM_KERNEL_1D(ht_scan_apply, HtScanParams)
static inline void launch_ht_tile_scan(const HtScanParams &p, cudaStream_t st) {
    launch_ht_scan_group(p, (u64)p.n_groups * p.nc, st);
    launch_ht_scan_top(p, p.nc, st);
    launch_ht_scan_apply(p, (u64)p.n_groups * p.nc, st);
}

// ---- per tile, in position order: PS / PL / PR
#if !defined(NLZM_EMU)
// shared table: in-tile offset + 1 of the bucket's last access (u16, 0 = not accessed in this tile yet, then the
// value comes from the table before the tile in global memory). Half the footprint of absolute positions
// => twice the resident warps of this latency-bound walk.
__global__ void __launch_bounds__(32) k_ht_prev(const HtTableParams p) {
    extern __shared__ __align__(16) u8 nlzm_smem[];
    u16 *tab = (u16 *)nlzm_smem;
    const u32 nc = 1u << p.c.bits, lane = threadIdx.x;
    const u32 *__restrict__ init = p.tile_last + (u64)blockIdx.x * nc;
    for (u32 i = lane; i < nc; i += 32) tab[i] = 0;
    const u64 t0 = p.pos0 + (u64)blockIdx.x * NLZM_HT_TILE;
    const u64 t1 = t0 + NLZM_HT_TILE < p.n_acc ? t0 + NLZM_HT_TILE : p.n_acc;
    const u32 t0p = (u32)t0;
    const u32 shift = 32 - p.c.bits;
    const u32 vmask = p.c.nbytes == 2 ? 0xFFFFu : 0xFFFFFFu;
    u32 *stage = (u32 *)(nlzm_smem + (size_t)nc * 2);      // NLZM_HT_STAGE + 8 text bytes of the walk, refilled every 64 steps
    __syncwarp();
    for (u64 base = t0; base < t1; base += 32) {
        const u32 in_stage = (u32)(base - t0) & (NLZM_HT_STAGE - 1);
        if (in_stage == 0) {
            // coalesced refill: the global-load latency is paid once per 2 KiB instead of once per step
            __syncwarp();
            const u64 *src = (const u64 *)(p.x + base);                  // base is a multiple of 32: aligned
            u64 *dst = (u64 *)stage;
            for (u32 i = lane; i < NLZM_HT_STAGE / 8 + 1; i += 32) dst[i] = (base + 8ull * i <= p.x_limit) ? src[i] : 0ull;
            __syncwarp();
        }
        const u64 a = base + lane;
        const bool live = a < t1;
        u32 b = 0xFFFFFFF0u;                                             // never equals a bucket or its neighbours
        if (live) {
            const u32 o = in_stage + lane;                               // byte offset inside the staged text
            const u32 w0 = stage[o >> 2], w1 = stage[(o >> 2) + 1];
            const u32 sh = (o & 3) * 8;
            const u32 v = sh ? ((w0 >> sh) | (w1 << (32 - sh))) : w0;
            b = ((v & vmask) * NLZM_HASH_MUL) >> shift;
        }
        u32 vs = 0, vl = 0, vr = 0;
        if (live) {
            u32 v = tab[b];
            vs = v ? t0p + v : init[b];
            if (p.c.rows == 2) {
                if (b > 0) { v = tab[b - 1]; vl = v ? t0p + v : init[b - 1]; }
                if (b + 1 < nc) { v = tab[b + 1]; vr = v ? t0p + v : init[b + 1]; }
            }
        }
        // The last access of each bucket among these 32 goes into the table (one writer per bucket); if no
        // lane shares a bucket and nobody else's access shows up in the cells a lane looks at, the 32
        // accesses do not interact (the common case) and the values read above are final.
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, b);
        __syncwarp();                                                    // all lanes have read the pre-step table
        const u32 me = (u32)(a - t0) + 1u;
        if (live && (peers >> lane) == 1u) tab[b] = (u16)me;
        __syncwarp();
        // Same bucket in a lower lane: the nearest one is the predecessor (exact, from the match mask).
        const unsigned lower = peers & ((1u << lane) - 1u);
        if (live && lower) vs = (u32)base + (31u - (u32)__clz((int)lower)) + 1u;
        // Neighbour buckets: the table now holds the LAST lane of each bucket touched in this step. If that lane
        // is below me it is my predecessor; if it is above me, a lower one may exist as well: rare, settled below.
        bool redo = false;
        if (live && p.c.rows == 2) {
            const u32 lo = (u32)(base - t0);
            if (b > 0) { const u32 t = tab[b - 1]; if (t > lo) { const u32 j = t - 1u - lo; if (j < lane) vl = (u32)base + j + 1u; else redo = true; } }
            if (b + 1 < nc) { const u32 t = tab[b + 1]; if (t > lo) { const u32 j = t - 1u - lo; if (j < lane) vr = (u32)base + j + 1u; else redo = true; } }
        }
        if (__any_sync(0xFFFFFFFFu, redo)) {
            // accesses inside these 32 positions, in order: a later lane overrides an earlier one
            for (u32 j = 0; j < 31; j++) {
                const u32 bj = __shfl_sync(0xFFFFFFFFu, b, j);
                if (j < lane) {
                    const u32 pj = (u32)(base + j) + 1u;
                    if (bj + 1 == b) vl = pj;
                    if (bj == b + 1) vr = pj;
                }
            }
        }
        if (live) {
            p.ps[a - p.pos0] = vs;
            if (p.c.rows == 2) { p.pl[a - p.pos0] = vl; p.pr[a - p.pos0] = vr; }
        }
        __syncwarp();
    }
}
static inline int launch_ht_prev(const HtTableParams &p, u64 grid, size_t smem, cudaStream_t st) {
    if (grid == 0) return 0;
    if (smem > 48 * 1024) {
        int e = nlzm_smem_opt_in((const void *)k_ht_prev, smem);
        if (e) return e;
    }
    nlzm_launch_begin("k_ht_prev", st);
    k_ht_prev<<<(unsigned)grid, 32, smem, st>>>(p);
    nlzm_launch_end(st);
    return (int)cudaGetLastError();
}
#else
// emulation: the same definition, one position after the other
static inline int launch_ht_prev(const HtTableParams &p, u64 grid, size_t, cudaStream_t st) {
    nlzm_launch_begin("k_ht_prev", st);
    const u32 nc = 1u << p.c.bits, shift = 32 - p.c.bits;
    std::vector<u32> tab(nc);
    for (u64 t = 0; t < grid; t++) {
        for (u32 i = 0; i < nc; i++) tab[i] = p.tile_last[t * nc + i];
        const u64 t0 = p.pos0 + t * NLZM_HT_TILE, t1 = t0 + NLZM_HT_TILE < p.n_acc ? t0 + NLZM_HT_TILE : p.n_acc;
        for (u64 a = t0; a < t1; a++) {
            const u32 b = ht_hash(p.x, a, p.c.nbytes) >> shift;
            p.ps[a - p.pos0] = tab[b];
            if (p.c.rows == 2) { p.pl[a - p.pos0] = b > 0 ? tab[b - 1] : 0; p.pr[a - p.pos0] = b + 1 < nc ? tab[b + 1] : 0; }
            tab[b] = (u32)a + 1u;
        }
    }
    nlzm_launch_end(st);
    return 0;
}
#endif

struct HtFindParams {
    const u8 *x;
    Geom g;
    HtCfg c;
    const u32 *ps, *pl, *pr;   // indexed by position - pos0
    u64 pos0;                  // first position PS/PL/PR exist for
    const u32 *coarse;         // [pos0 >> coarse_log][1 << bits]: table before each coarse tile of the far prefix
    const u32 *coarse_first;   // same shape: first access (+1) inside the coarse tile, 0 = none
    const u32 *coarse_last;    // same shape: last access (+1) inside the coarse tile, 0 = none
    const u32 *coarse_count;   // same shape: number of accesses inside the coarse tile
    u32 coarse_log;
    u32 warp_scan;             // 1: all lanes of a warp run the same chain (snapshot kernel) and scan together
    const u32 *snap;           // raw content of every cell at time pos0 (null when pos0 == 0): a chain that would leave
                               // [pos0, ..) stops here instead of walking the far prefix
    u64 own_b;
    u32 bt_on;           // exhaustive BT4 runs too: it reports a candidate at least as near and as long for every
                         // match of 4+ bytes, so those need not be queued twice (they would be merged away)
    TupleSink sink;
};

// entry an access at position w stores: full shifted position OR-ed with the check bits (NLZM.cpp:913)
DEV u32 ht_entry(const HtFindParams &p, u64 w) {
    const u32 cmask = (1u << (32 - p.g.hb)) - 1;
    return geom_P(p.g, w) | ((ht_hash(p.x, w, p.c.nbytes) & cmask) << p.g.hb);
}

// last access (+1, 0 = none) before position q in `bucket`, for q in the far prefix: scan back to the start
// of q's coarse tile, then the coarse table (which holds the last access before that tile)
DEV u32 ht_last_before(const HtFindParams &p, u32 bucket, u64 q) {
    const u64 tile = q >> p.coarse_log, t0 = tile << p.coarse_log;
    const u64 cell = tile * ((u64)1 << p.c.bits) + bucket;
    const u32 first = p.coarse_first[cell];
    if (first == 0 || (u64)first - 1 >= q) return p.coarse[cell];      // no access of this bucket in [t0, q)
    const u32 last = p.coarse_last[cell];
    if ((u64)last - 1 < q) return last;                                 // the tile's last access lies before q
    if (p.coarse_count[cell] == 2) return first;                        // first < q <= last and nothing in between
    const u32 shift = 32 - p.c.bits;                                    // three or more accesses around q: scan back
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__)
    if (p.warp_scan) {
        // snapshot kernel: the 32 lanes of a warp follow the same chain, so the scan is theirs together —
        // 32 positions per step, coalesced, the nearest hit picked with a ballot
        const u32 lane = threadIdx.x & 31;
        for (u64 hi = q; hi > t0; hi = hi >= 32 ? hi - 32 : 0) {
            const u64 a = hi - 1 - lane;                                    // hi-1, hi-2, ... (wraps below 0: masked)
            const bool in = hi > lane && a >= t0;
            const bool hit = in && (ht_hash(p.x, a, p.c.nbytes) >> shift) == bucket;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
            if (m) return (u32)(hi - 1 - (u32)(__ffs((int)m) - 1)) + 1u;
            if (hi < 32) break;
        }
        return p.coarse[cell];
    }
#endif
    for (u64 a = q; a-- > t0; )
        if ((ht_hash(p.x, a, p.c.nbytes) >> shift) == bucket) return (u32)a + 1u;
    return p.coarse[cell];
}
DEV u32 ht_ps(const HtFindParams &p, u64 q, u32 bucket) { return q >= p.pos0 ? p.ps[q - p.pos0] : ht_last_before(p, bucket, q); }
DEV u32 ht_pl(const HtFindParams &p, u64 q, u32 bucket) { return q >= p.pos0 ? p.pl[q - p.pos0] : ht_last_before(p, bucket - 1, q); }

// raw u32 content of cell `cell` as seen by a reader at time t, given the last accesses (+1) before t
// of bucket `cell` (w0: writes its own entry) and of bucket `cell - 1` (w1: pushes cell-1's old content)
DEV u32 ht_cell_value(const HtFindParams &p, u32 cell, u64 t, u32 w0, u32 w1) {
    while (true) {
        if (w0 == 0 && w1 == 0) return NLZM_NONE32;                         // never written
        if (cell == 0) {                                                     // only bucket 0 writes cell 0 ...
            const u64 w = w0 - 1;
            if (geom_epoch(p.g, w) != geom_epoch(p.g, t)) return NLZM_NONE32;  // ... and every ring shift clears it
            return ht_entry(p, w);
        }
        // nothing wrote this cell since pos0: it still holds what it held then
        if (p.snap && t >= p.pos0 && (u64)w0 <= p.pos0 && (u64)w1 <= p.pos0) return p.snap[cell];
        if (w0 > w1) return ht_entry(p, w0 - 1);
        const u64 q = w1 - 1;                 // bucket cell-1 accessed at q and pushed the old content of cell-1
        t = q;
        cell -= 1;                            // = q's bucket
        w0 = ht_ps(p, q, cell);
        w1 = (p.c.rows == 2 && cell > 0) ? ht_pl(p, q, cell) : 0u;
    }
}

// content of every cell at time pos0, from the coarse tables of the far prefix (one thread per cell; the only
// place where chains walk the far prefix)
struct HtSnapParams { HtFindParams f; const u32 *row_at_pos0; u32 *snap_out; };
DEV void ht_snapshot_body(const HtSnapParams &p, u64 i) {
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__)
    const u32 cell = (u32)(i >> 5);                                      // one warp per cell
#else
    const u32 cell = (u32)i;
#endif
    const u32 nc = 1u << p.f.c.bits;
    HtFindParams f = p.f;
    f.snap = nullptr;
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__)
    f.warp_scan = 1;
#endif
    const u32 w0 = cell < nc ? p.row_at_pos0[cell] : 0u;
    const u32 w1 = (f.c.rows == 2 && cell > 0) ? p.row_at_pos0[cell - 1] : 0u;
    const u32 v = ht_cell_value(f, cell, f.pos0, w0, w1);
#if !defined(NLZM_EMU) && defined(__CUDA_ARCH__)
    if ((threadIdx.x & 31) == 0)
#endif
    p.snap_out[cell] = v;
}
NLZM_KERNEL_1D(ht_snapshot, HtSnapParams)

DEV void ht_find_body(const HtFindParams &p, u64 i) {
    const u64 a = p.own_b + i;
    const u32 hash = ht_hash(p.x, a, p.c.nbytes);
    const u32 cmask = (1u << (32 - p.g.hb)) - 1;
    const u32 chk = hash & cmask;
    const u32 b = hash >> (32 - p.c.bits);
    const u32 P = geom_P(p.g, a);
    const u32 rem = geom_rem(p.g, a);
    const u32 cap = rem < NLZM_MATCH_MAX ? rem : NLZM_MATCH_MAX;      // NLZM.cpp:915
    const u64 base = a - P;
    const u64 ai = a - p.pos0;
    const u32 ps = p.ps[ai];
    u32 best = 1;                                                     // MATCH_MIN - 1, NLZM.cpp:917
    for (u32 r = 0; r < p.c.rows; r++) {
        const u32 row = r == 0 ? ht_cell_value(p, b, a, ps, (p.c.rows == 2 && b > 0) ? p.pl[ai] : 0u)
                               : ht_cell_value(p, b + 1, a, p.pr[ai], ps);
        if (best < cap && (row >> p.g.hb) == chk) {
            const u32 sp = row & (p.g.W - 1);
            if (sp < P && P - sp <= p.g.W - 1) {
                const u32 m = lcp_cap(p.x, base + sp, a, cap);
                if (m > best && m >= match_min(P - sp)) {
                    if (m < 4 || !p.bt_on) tuple_append(p.sink, (u32)i, P - sp, m);
                    best = m;
                }
            }
        }
    }
}
NLZM_KERNEL_1D(ht_find, HtFindParams)

// bucket sort helper kept for bt_short.cuh (small windows): inverse permutation of a sorted value list
struct HtInvParams { const u32 *vals; u32 *inv; u32 first_ev; };
DEV void ht_inv_body(const HtInvParams &p, u64 j) {
    const u32 v = p.vals[j];
    if (v >= p.first_ev) p.inv[v - p.first_ev] = (u32)j;
}
NLZM_KERNEL_1D(ht_inv, HtInvParams)
