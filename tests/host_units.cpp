// Unit checks of the host pipeline's building blocks (compiled and run by tests/test_host_units.py).
//   1. Staircase::merge_steps == one Update per step, for random step lists that satisfy the ABI's
//      contract (strictly increasing len and dist), on top of random carried state, with advance().
//   2. FrameWriter -> FrameReader round trip of random symbols on adapting tables mixed with raw bits.
//   3. PriceList: non-increasing in the frequency, 0 for certainty, 32 units per halving (truncated).
//   4. split_distance / join_distance are inverse; shortest_len thresholds.
#include "../nlzm_b200/csrc/host/frame_coder.hpp"
#include "../nlzm_b200/csrc/host/parser.hpp"
#include <random>
#include <stdio.h>
#include <stdlib.h>

using namespace nlzm_host;

#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static void staircase_merge_equals_updates() {
    std::mt19937 rng(1);
    for (int round = 0; round < 20000; round++) {
        Staircase a, b;
        int positions = 1 + rng() % 6;
        for (int pos = 0; pos < positions; pos++) {
            uint32_t n = rng() % 6, len = 1, dist = 0;
            uint32_t dists[8], lens[8];
            uint32_t k = 0;
            for (uint32_t j = 0; j < n; j++) {
                len += 1 + rng() % (j == 0 ? 12 : 60);
                dist += 1 + rng() % 5000;
                if (len > kLenMax) break;
                dists[k] = dist; lens[k] = len; k++;
            }
            for (uint32_t j = 0; j < k; j++) a.Update(dists[j], (uint16_t)lens[j]);
            b.merge_steps(k, [&](uint32_t j) { return candidate_entry(dists[j], candidate_code(dists[j])); }, [&](uint32_t j) { return lens[j]; });
            CHECK(a.top == b.top);
            for (uint32_t l = 1; l <= a.top; l++) CHECK(a.entry(l) == b.entry(l));
            int moves = rng() % 4;
            for (int m = 0; m < moves; m++) { a.advance(); b.advance(); }
            CHECK(a.top == b.top);
            for (uint32_t l = 1; l <= a.top; l++) CHECK(a.entry(l) == b.entry(l));
        }
    }
    // the sliding buffer wraps (16384 advances) without losing the carried entries
    Staircase s;
    for (uint32_t i = 0; i < 40000; i++) {
        s.Update(7 + i % 3, 200);
        uint32_t d = s[150];
        s.advance();
        CHECK(s.top == 199 && s[149] == d);
    }
}

static void frame_round_trip() {
    std::mt19937 rng(2);
    for (int round = 0; round < 200; round++) {
        Table<2> t2w, t2r; Table<3> t3w, t3r; Table<4> t4w, t4r;
        t2w.reset(); t2r.reset(); t3w.reset(); t3r.reset(); t4w.reset(); t4r.reset();
        FrameWriter w;
        w.begin();
        struct Op { int kind; uint32_t v, nb; };
        std::vector<Op> ops;
        int n = rng() % 3000;
        int skew = 1 + rng() % 4;
        for (int i = 0; i < n; i++) {
            int kind = rng() % 4;
            uint32_t r = rng();
            for (int s = 1; s < skew; s++) r &= rng();          // skewed symbols make the tables adapt hard
            if (kind == 0) { int y = r & 3; w.put(t2w, y); t2w.adapt(y); ops.push_back({0, (uint32_t)y, 0}); }
            if (kind == 1) { int y = r & 7; w.put(t3w, y); t3w.adapt(y); ops.push_back({1, (uint32_t)y, 0}); }
            if (kind == 2) { int y = r & 15; w.put(t4w, y); t4w.adapt(y); ops.push_back({2, (uint32_t)y, 0}); }
            if (kind == 3) { uint32_t nb = 1 + rng() % 22; uint32_t v = rng() & ((1u << nb) - 1); w.put_raw(v, nb); ops.push_back({3, v, nb}); }
        }
        std::vector<uint8_t> buf(5, 0xAA);                        // frames append to what is there
        size_t size = w.end(buf);
        CHECK(size + 5 == buf.size());
        buf.insert(buf.end(), 4, 0);
        FrameReader r;
        if (ops.empty()) { CHECK(r.begin(buf.data() + 5, buf.size() - 5) == 0); continue; }   // ops == 0 reads as the end marker
        CHECK(r.begin(buf.data() + 5, buf.size() - 5) == (int64_t)size);
        CHECK(r.ops_left() == ops.size());
        for (const Op &o : ops) {
            if (o.kind == 0) { int y = r.get(t2r); CHECK((uint32_t)y == o.v); t2r.adapt(y); }
            if (o.kind == 1) { int y = r.get(t3r); CHECK((uint32_t)y == o.v); t3r.adapt(y); }
            if (o.kind == 2) { int y = r.get(t4r); CHECK((uint32_t)y == o.v); t4r.adapt(y); }
            if (o.kind == 3) CHECK(r.get_raw(o.nb) == o.v);
            CHECK(!r.bad());
        }
        CHECK(r.ops_left() == 0);
        // every slice stays non-empty however hard a table is pushed
        for (int y = 0; y < 16; y++) CHECK(t4w.freq(y) > 0);
        for (int y = 0; y < 8; y++) CHECK(t3w.freq(y) > 0);
        for (int y = 0; y < 4; y++) CHECK(t2w.freq(y) > 0);
    }
    Table<4> t; t.reset();
    for (int i = 0; i < 5000; i++) t.adapt(0);
    for (int y = 0; y < 16; y++) CHECK(t.freq(y) > 0);
    for (int i = 0; i < 5000; i++) t.adapt(15);
    for (int y = 0; y < 16; y++) CHECK(t.freq(y) > 0);
    CHECK(t.cum[0] == 0 && t.cum[16] == kProbOne);
}

static void price_list() {
    const PriceList &p = prices();
    for (uint32_t f = 64; f < kProbOne; f += 64) CHECK(p(f) >= p(f + 63));
    CHECK(p(kProbOne - 1) == 0);                     // near certainty costs nothing
    // 32 units per bit, truncated: exact halvings land one unit below the multiple of 32
    CHECK(p(kProbOne / 2) == 31 && p(kProbOne / 4) == 63 && p(kProbOne / 16) == 127 && p(128) == 223 && p(64) == 255);
}

static void distance_codes() {
    std::mt19937 rng(3);
    for (int i = 0; i < 200000; i++) {
        uint32_t dist = 1 + (rng() >> (4 + rng() % 28));
        if (dist > (1u << 28)) continue;
        DistCode dc = split_distance(dist);
        CHECK(dc.slot < 56);
        uint32_t back = dc.slot < 4 ? dc.slot : join_distance(dc.slot, dc.raw_bits, dc.raw);
        CHECK(back + 1 == dist);
        if (dc.slot >= 4) CHECK(dc.raw_bits == (dc.slot >> 1) - 1);
    }
    CHECK(shortest_len(1) == 2 && shortest_len(255) == 2 && shortest_len(256) == 3 && shortest_len(4095) == 3);
    CHECK(shortest_len(4096) == 4 && shortest_len((1u << 20) - 1) == 4 && shortest_len(1u << 20) == 5);
}

int main() {
    staircase_merge_equals_updates();
    frame_round_trip();
    price_list();
    distance_codes();
    printf("host units ok\n");
    return 0;
}
