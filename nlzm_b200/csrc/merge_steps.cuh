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

// ================================================================================================
// Stage M, bin form. The tuples are radix-sorted by BIN only (a bin = 512 consecutive positions: half the
// key bits, half the sort passes); one CTA then finishes a bin in shared memory: tuples are grouped by
// position with a counting sort, every position's handful of candidates is ordered and filtered by its own
// thread, and the surviving steps leave as final 6-byte records, bin-contiguous in a staging array. A
// global scan of the per-position counts gives the CSR offsets; k_bin_place moves each bin's records to
// their place (one contiguous copy per bin).
// ================================================================================================
#define NLZM_BIN_LOG 9u
#define NLZM_BIN (1u << NLZM_BIN_LOG)
#ifdef NLZM_EMU
#define NLZM_BIN_THREADS 16u                    // the emulation pays for every barrier with host threads
#else
#define NLZM_BIN_THREADS 256u
#endif
#define NLZM_BIN_CAP 4096u                      // tuples of one pass through shared memory
#define NLZM_BIN_SMEM (NLZM_BIN_CAP * 8 + NLZM_BIN * 12 + (NLZM_BIN_THREADS + 40) * 4)

struct BinBoundsParams { const u64 *keys; u32 nt; u32 *bin_start; };
DEV void bin_bounds_body(const BinBoundsParams &p, u64 j) {        // j in [0, n_bins]: first tuple of bin j
    const u64 want = j << (9 + NLZM_BIN_LOG);
    u32 lo = 0, hi = p.nt;
    while (lo < hi) { const u32 mid = lo + ((hi - lo) >> 1); if (p.keys[mid] < want) lo = mid + 1; else hi = mid; }
    p.bin_start[j] = lo;
}
NLZM_KERNEL_1D(bin_bounds, BinBoundsParams)

struct BinFinishParams {
    const u64 *keys; const u32 *dist;     // tuples sorted by bin
    const u32 *bin_start;                 // n_bins + 1
    u32 n_own;
    u32 *count;                           // out: surviving steps per position (n_own + 1, zeroed)
    Step *staging;                        // out: surviving steps, bin b at staging[bin_start[b] ...)
};

// exclusive scan of v[0..NLZM_BIN) in place by the whole CTA (NLZM_BIN / threads consecutive entries per thread,
// thread totals combined through `part`); returns the grand total
DEV u32 bin_scan(u32 *v, u32 *part, u32 tid) {
    const u32 per = NLZM_BIN / NLZM_BIN_THREADS;
    u32 sum = 0;
    for (u32 i = 0; i < per; i++) sum += v[tid * per + i];
    part[tid] = sum;
    NLZM_CTA_SYNC();
    if (tid < 32) {                                           // one warp scans the thread totals
        const u32 chunk = NLZM_BIN_THREADS / 32 ? NLZM_BIN_THREADS / 32 : 1;
        if (tid * chunk < NLZM_BIN_THREADS) {
            u32 run = 0;
            for (u32 i = 0; i < chunk && tid * chunk + i < NLZM_BIN_THREADS; i++) { const u32 t = part[tid * chunk + i]; part[tid * chunk + i] = run; run += t; }
            part[NLZM_BIN_THREADS + tid] = run;
        }
    }
    NLZM_CTA_SYNC();
    if (tid == 0) {
        u32 run = 0;
        const u32 groups = NLZM_BIN_THREADS < 32 ? NLZM_BIN_THREADS : 32;
        for (u32 g = 0; g < groups; g++) { const u32 t = part[NLZM_BIN_THREADS + g]; part[NLZM_BIN_THREADS + g] = run; run += t; }
        part[NLZM_BIN_THREADS + 32] = run;
    }
    NLZM_CTA_SYNC();
    {
        const u32 chunk = NLZM_BIN_THREADS / 32 ? NLZM_BIN_THREADS / 32 : 1;
        u32 run = part[tid] + part[NLZM_BIN_THREADS + tid / chunk];
        for (u32 i = 0; i < per; i++) { const u32 t = v[tid * per + i]; v[tid * per + i] = run; run += t; }
    }
    NLZM_CTA_SYNC();
    return part[NLZM_BIN_THREADS + 32];
}

// order one position's candidates by (length descending, distance ascending) and keep the lower envelope:
// a candidate survives iff it is nearer than everything at least as long. Survivors end up packed at the
// front, longest first. g[] holds len | dist << 9.
DEV u32 bin_envelope(u64 *g, u32 n) {
    for (u32 i = 1; i < n; i++) {                             // insertion sort, a handful of elements
        const u64 v = g[i];
        const u64 kv = ((u64)(511u - (u32)(v & 511u)) << 40) | (v >> 9);
        u32 j = i;
        while (j > 0) {
            const u64 w = g[j - 1];
            const u64 kw = ((u64)(511u - (u32)(w & 511u)) << 40) | (w >> 9);
            if (kw <= kv) break;
            g[j] = w;
            --j;
        }
        g[j] = v;
    }
    u32 best = 0xFFFFFFFFu, k = 0;
    for (u32 i = 0; i < n; i++) {
        const u32 d = (u32)(g[i] >> 9);
        if (d < best) { best = d; g[k++] = g[i]; }
    }
    return k;
}

DEV void bin_finish_cta(const BinFinishParams &p, u32 bid, u32 tid, u8 *smem) {
    u64 *tup = (u64 *)smem;                                   // len | dist << 9 of the tuples of this pass, grouped by position
    u32 *cnt = (u32 *)(tup + NLZM_BIN_CAP);                   // tuples per position, then fill counters
    u32 *off = cnt + NLZM_BIN;                                // start of a position's group inside tup[]
    u32 *kept = off + NLZM_BIN;                               // surviving steps per position, then their bin-local offsets
    u32 *part = kept + NLZM_BIN;                              // scan scratch: NLZM_BIN_THREADS + 33 words
    const u32 base = p.bin_start[bid], m = p.bin_start[bid + 1] - base;
    const u32 pos0 = bid << NLZM_BIN_LOG;
    const u32 n_pos = (p.n_own - pos0) < NLZM_BIN ? (p.n_own - pos0) : NLZM_BIN;
    if (m == 0) return;
    u32 lo = 0, written = 0;
    while (lo < n_pos) {
        for (u32 i = tid; i < NLZM_BIN; i += NLZM_BIN_THREADS) { cnt[i] = 0; kept[i] = 0; }
        NLZM_CTA_SYNC();
        for (u32 i = tid; i < m; i += NLZM_BIN_THREADS) {
            const u32 q = (u32)(p.keys[base + i] >> 9) & (NLZM_BIN - 1);
            if (q >= lo) nlzm_atomic_add(cnt + q, 1u);
        }
        NLZM_CTA_SYNC();
        for (u32 i = tid; i < NLZM_BIN; i += NLZM_BIN_THREADS) off[i] = cnt[i];
        NLZM_CTA_SYNC();
        const u32 total = bin_scan(off, part, tid);           // off[q] = tuples of positions [lo, q)
        // positions [lo, hi) whose tuples fit into one pass: all of them in the common case
        u32 hi = n_pos;
        if (total > NLZM_BIN_CAP) {
            u32 a = lo + 1, b = n_pos;                        // largest hi with off[hi] <= CAP (off is non-decreasing)
            while (a < b) { const u32 mid = (a + b + 1) >> 1; if ((mid < NLZM_BIN ? off[mid] : total) <= NLZM_BIN_CAP) a = mid; else b = mid - 1; }
            hi = a;
        }
        for (u32 i = tid; i < NLZM_BIN; i += NLZM_BIN_THREADS) cnt[i] = 0;
        NLZM_CTA_SYNC();
        for (u32 i = tid; i < m; i += NLZM_BIN_THREADS) {
            const u64 k = p.keys[base + i];
            const u32 q = (u32)(k >> 9) & (NLZM_BIN - 1);
            if (q >= lo && q < hi) {
                const u32 slot = off[q] + nlzm_atomic_add(cnt + q, 1u);
                if (slot < NLZM_BIN_CAP) tup[slot] = (k & 511u) | ((u64)p.dist[base + i] << 9);
            }
        }
        NLZM_CTA_SYNC();
        for (u32 q = lo + tid; q < hi; q += NLZM_BIN_THREADS) {
            u32 n = cnt[q];
            if (off[q] + n > NLZM_BIN_CAP) n = off[q] < NLZM_BIN_CAP ? NLZM_BIN_CAP - off[q] : 0;   // a single position beyond a whole pass: cannot happen
            const u32 k = bin_envelope(tup + off[q], n);
            kept[q] = k;
            p.count[pos0 + q] = k;
        }
        NLZM_CTA_SYNC();
        const u32 kept_total = bin_scan(kept, part, tid);     // bin-local offsets of the survivors, positions in order
        for (u32 q = lo + tid; q < hi; q += NLZM_BIN_THREADS) {
            const u64 *g = tup + off[q];
            const u32 k = p.count[pos0 + q];
            Step *out = p.staging + base + written + kept[q];
            for (u32 i = 0; i < k; i++) {                     // ascending length = reverse of the kept order
                const u64 v = g[k - 1 - i];
                const u32 d = (u32)(v >> 9);
                Step s;
                s.dist_lo = (u16)d;
                s.dist_hi = (u16)((d >> 16) | ((match_min(d) - 2u) << 12));          // bits 12..13: shortest length - 2
                s.len = (u16)((u32)(v & 511u) | (step_dist_slot(d) << 9));            // bits 9..14: distance slot
                out[i] = s;
            }
        }
        NLZM_CTA_SYNC();
        written += kept_total;
        lo = hi;
    }
}
NLZM_KERNEL_CTA_OCC(bin_finish, BinFinishParams, NLZM_BIN_THREADS, 4)

// bin b's records: staging[bin_start[b] ...) -> steps[offsets[first position of b] ...)
struct BinPlaceParams { const Step *staging; const u32 *bin_start; const u32 *offsets; u32 n_own; Step *steps; };
DEV void bin_place_cta(const BinPlaceParams &p, u32 bid, u32 tid, u8 *) {
    const u32 pos0 = bid << NLZM_BIN_LOG;
    const u32 pos1 = (pos0 + NLZM_BIN) < p.n_own ? (pos0 + NLZM_BIN) : p.n_own;
    const u32 o0 = p.offsets[pos0], k = p.offsets[pos1] - o0;
    const u16 *src = (const u16 *)(p.staging + p.bin_start[bid]);
    u16 *dst = (u16 *)(p.steps + o0);
    for (u32 i = tid; i < k * 3; i += NLZM_BIN_THREADS) dst[i] = src[i];
}
NLZM_KERNEL_CTA(bin_place, BinPlaceParams, NLZM_BIN_THREADS)
