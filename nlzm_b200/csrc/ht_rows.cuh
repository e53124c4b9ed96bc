// Stage H — HT2 / HT3 short-match finders (NLZM.cpp:893-957) without a table.
//
// The reference keeps `rows << bits` u32 cells; an access with bucket b reads cells b .. b+rows-1
// (rows overlap because the row pointer is `rows + bucket`, NLZM.cpp:912), then writes its own
// entry into cell b and pushes the old content of cell b into cell b+1. The content a reader sees
// is therefore "the last write before me": every access is turned into `rows` write events
// (cell, time), the events are radix-sorted by cell (stable => time order inside a cell), and a
// reader takes the event just before its own in the cell's list. A row-1 event carries the old
// content of the cell to its left at that time, which is resolved the same way (a short chain).
// MatchFinderHT::Shift as written only clears cell 0 at each ring shift (NLZM.cpp:940-957).
#pragma once
#include "common.cuh"
#include "dc_levels.cuh"

struct HtCfg {
    u32 rows;      // 1 (HT2) or 2 (HT3)
    u32 bits;      // bucket index bits
    u32 nbytes;    // 2 or 3 hashed bytes
};

HD u32 ht_hash(const u8 *__restrict__ x, u64 a, u32 nbytes) {
    u32 v = load4(x, a) & (nbytes == 2 ? 0xFFFFu : 0xFFFFFFu);      // VALUE2 / VALUE3, NLZM.cpp:741-742
    return v * NLZM_HASH_MUL;
}

struct HtEventParams { const u8 *x; HtCfg c; u32 *keys; u32 *vals; };
DEV void ht_event_body(const HtEventParams &p, u64 a) {
    u32 b = ht_hash(p.x, a, p.c.nbytes) >> (32 - p.c.bits);
    for (u32 r = 0; r < p.c.rows; r++) {
        p.keys[a * p.c.rows + r] = b + r;            // cell written by this access at row r
        p.vals[a * p.c.rows + r] = (u32)(a * p.c.rows + r);
    }
}
NLZM_KERNEL_1D(ht_event, HtEventParams)

// inverse permutation, only for events at or after `first_ev` (the range being answered): tables
// never age, so the sorted event list covers the whole prefix, but only own events start a lookup
struct HtInvParams { const u32 *vals; u32 *inv; u32 first_ev; };
DEV void ht_inv_body(const HtInvParams &p, u64 j) {
    const u32 v = p.vals[j];
    if (v >= p.first_ev) p.inv[v - p.first_ev] = (u32)j;
}
NLZM_KERNEL_1D(ht_inv, HtInvParams)

struct HtFindParams {
    const u8 *x;
    Geom g;
    HtCfg c;
    const u32 *skeys;    // sorted cells
    const u32 *svals;    // event ids in (cell, time) order
    const u32 *inv;      // event id - first_ev -> index in the sorted arrays
    u32 first_ev;        // first event id covered by inv
    u32 n_ev;            // number of sorted events
    u64 own_b;
    TupleSink sink;
};

// index of event `ev` (which writes `cell`) in the sorted arrays
DEV u32 ht_event_index(const HtFindParams &p, u32 ev, u32 cell) {
    if (ev >= p.first_ev) return p.inv[ev - p.first_ev];
    // an event before the answered range (rare: a chain that reaches back): binary search by (cell, id)
    u32 lo = 0, hi = p.n_ev;
    while (lo < hi) {
        const u32 mid = lo + ((hi - lo) >> 1);
        const u32 c = p.skeys[mid];
        if (c < cell || (c == cell && p.svals[mid] < ev)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// raw u32 content of the cell that event `ev` (of the access at time t) is about to overwrite
DEV u32 ht_cell_before(const HtFindParams &p, u32 ev, u64 t, u32 cell0) {
    const u32 cmask = (1u << (32 - p.g.hb)) - 1;
    u32 j = ht_event_index(p, ev, cell0);
    while (true) {
        if (j == 0) return NLZM_NONE32;
        const u32 cell = p.skeys[j];
        if (p.skeys[j - 1] != cell) return NLZM_NONE32;          // nobody wrote this cell before
        const u32 pe = p.svals[j - 1];
        const u64 w = pe / p.c.rows;                            // the writer's position
        const u32 kind = pe - (u32)w * p.c.rows;
        if (cell == 0 && geom_epoch(p.g, w) != geom_epoch(p.g, t)) return NLZM_NONE32;   // cleared by a ring shift
        if (kind == 0) {
            // the writer stored its own entry: full shifted position OR-ed with the check bits
            return geom_P(p.g, w) | ((ht_hash(p.x, w, p.c.nbytes) & cmask) << p.g.hb);   // NLZM.cpp:913
        }
        // the writer pushed the old content of its own bucket cell (cell - 1) here
        t = w;
        j = ht_event_index(p, (u32)w * p.c.rows, cell - 1);
    }
}

DEV void ht_find_body(const HtFindParams &p, u64 i) {
    const u64 a = p.own_b + i;
    const u32 hash = ht_hash(p.x, a, p.c.nbytes);
    const u32 cmask = (1u << (32 - p.g.hb)) - 1;
    const u32 chk = hash & cmask;
    const u32 P = geom_P(p.g, a);
    const u32 rem = geom_rem(p.g, a);
    const u32 cap = rem < NLZM_MATCH_MAX ? rem : NLZM_MATCH_MAX;      // NLZM.cpp:915
    const u64 base = a - P;
    u32 best = 1;                                                     // MATCH_MIN - 1, NLZM.cpp:917
    for (u32 r = 0; r < p.c.rows; r++) {
        const u32 row = ht_cell_before(p, (u32)a * p.c.rows + r, a, (hash >> (32 - p.c.bits)) + r);
        if (best < cap && (row >> p.g.hb) == chk) {
            const u32 sp = row & (p.g.W - 1);
            if (sp < P && P - sp <= p.g.W - 1) {
                const u32 m = lcp_cap(p.x, base + sp, a, cap);
                if (m > best && m >= match_min(P - sp)) {
                    tuple_append(p.sink, (u32)i, P - sp, m);
                    best = m;
                }
            }
        }
    }
}
NLZM_KERNEL_1D(ht_find, HtFindParams)
