// stream_model.hpp — the adaptive nibble model of the NLZM stream and its integer price list.
//
// Host side of the engine boundary. Restates, in this project's own terms:
//   cumulative-frequency tables, their initial state and adaptation     NLZM.cpp:212-381
//   the price of a symbol in 1/32 bit (log2 table)                      NLZM.cpp:96-124, 435-438
//   the model's table set and the repeat-distance memory                NLZM.cpp:1124-1205
//   prices of literal / match / repeat commands                         NLZM.cpp:1208-1272, 1418-1426
//   the minimum length a distance may be coded with                     NLZM.cpp:826-834
// All arithmetic is integer and must agree bit for bit: the parser's decisions, and therefore the
// stream, depend on every price.
#ifndef NLZM_HOST_STREAM_MODEL_HPP
#define NLZM_HOST_STREAM_MODEL_HPP

#include <stdint.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace nlzm_host {

constexpr int kProbBits = 14;                    // probabilities are 14-bit fixed point
constexpr uint32_t kProbOne = 1u << kProbBits;
constexpr int kAdaptShift = 7;                   // adaptation rate 2^-7
constexpr int kPriceShift = 5;                   // prices are in 1/32 bit
constexpr uint32_t kLenMin = 2, kLenMax = kLenMin + 255 + 7;

enum Command : uint8_t { kLiteral = 0, kMatch = 1, kRepeat = 2, kNone = 0xFF };

static inline uint32_t bit_length(uint32_t v) { return 32u - (uint32_t)__builtin_clz(v); }

// shortest length a match at this distance may have: 2, +1 past 2^8, 2^12, 2^20
static inline uint32_t shortest_len(uint32_t dist) {
    return kLenMin + (dist >= (1u << 8)) + (dist >= (1u << 12)) + (dist >= (1u << 20));
}

// price[f >> 6] = 32 * log2(2^14 / f), from 32 rounds of fixed-point multiply-and-renormalise: the
// number of renormalisation shifts of (i/256)^32 is the integer price
struct PriceList {
    uint16_t of_bucket[256];
    PriceList() {
        for (uint32_t i = 1; i < 256; i++) {
            uint32_t mant = 1u << 16;
            uint16_t shifts = 0;
            for (int round = 0; round < (1 << kPriceShift); round++) {
                uint32_t prod = (i * mant) >> 8;
                uint32_t up = 16u - bit_length(prod);        // shift that brings prod back to [2^15, 2^16)
                shifts = (uint16_t)(shifts + up);
                mant = prod << up;
            }
            of_bucket[i] = shifts;
        }
        of_bucket[0] = of_bucket[1];
    }
    uint32_t operator()(uint32_t freq) const { return of_bucket[freq >> (kProbBits - 8)]; }
};

inline const PriceList &prices() {
    static const PriceList p;
    return p;
}

// cumulative table over 2^BITS symbols: cum[y]..cum[y+1] is symbol y's slice of [0, 2^14)
template <int BITS> struct Table {
    static constexpr int N = 1 << BITS;
    uint16_t cum[N + 1];

    void reset() {
        for (int i = 0; i <= N; i++) cum[i] = (uint16_t)(i * (int)(kProbOne >> BITS));
    }
    uint32_t low(int y) const { return cum[y]; }
    uint32_t freq(int y) const { return (uint32_t)cum[y + 1] - cum[y]; }
    uint32_t price(int y) const { return prices()(freq(y)); }

    // after coding y: boundaries at or below y drift towards their floor (their own index, which
    // keeps every slice non-empty), boundaries above y towards a ceiling just past 2^14. All values
    // and differences fit in 16 signed bits, so the whole row updates in 16-bit lanes.
    struct Targets {
        int16_t row[N][N];
        constexpr Targets() : row{} {
            constexpr int ceiling_bias = (int)kProbOne + (1 << kAdaptShift) - 1 - N;
            for (int y = 0; y < N; y++)
                for (int x = 0; x < N; x++) row[y][x] = (int16_t)(x <= y ? x : ceiling_bias + x);
        }
    };
    static constexpr Targets kTargets{};
    void adapt(int y) {
        const int16_t *t = kTargets.row[y];
        for (int x = 0; x < N; x++) cum[x] = (uint16_t)(cum[x] + ((int16_t)(t[x] - (int16_t)cum[x]) >> kAdaptShift));
    }
    // symbol whose slice holds f (decoder): the number of boundaries cum[1..N-1] that are <= f
    int find(uint32_t f) const {
#if defined(__SSE2__)
        if (N >= 8) {                                    // one compare per 8 boundaries; cum[N] = 2^14 > f ends the scan
            const __m128i vf = _mm_set1_epi16((short)f);
            __m128i above = _mm_cmpgt_epi16(_mm_loadu_si128((const __m128i *)(cum + 1)), vf);
            if (N == 8) return __builtin_ctz((unsigned)_mm_movemask_epi8(above)) >> 1;
            above = _mm_packs_epi16(above, _mm_cmpgt_epi16(_mm_loadu_si128((const __m128i *)(cum + 9)), vf));
            return __builtin_ctz((unsigned)_mm_movemask_epi8(above));
        }
#endif
        int y = 0;
        for (int half = N >> 1; half; half >>= 1)
            if (f >= cum[y + half]) y += half;
        return y;
    }
};

// the last four distinct distances; a hit does not reorder (NLZM.cpp:1154-1181)
struct RecentDistances {
    uint32_t d[4];
    void reset() { d[0] = 1; d[1] = 2; d[2] = 3; d[3] = 4; }
    int index_of(uint32_t dist) const {
        for (int i = 0; i < 4; i++)
            if (d[i] == dist) return i;
        return -1;
    }
    void remember(uint32_t dist) {
        if (index_of(dist) >= 0) return;
        d[3] = d[2]; d[2] = d[1]; d[1] = d[0]; d[0] = dist;
    }
};

// distance -> (slot, number of raw bits, raw bits): slots 0..3 are distances 1..4; above that the
// slot is 2*(bit length - 1) + (second highest bit) of dist-1 and the rest is sent raw
struct DistCode {
    uint32_t slot, raw_bits, raw;
};
static inline DistCode split_distance(uint32_t dist) {
    uint32_t v = dist - 1;
    if (v < 4) return {v, 0, 0};
    uint32_t nb = bit_length(v);
    uint32_t raw_bits = nb - 2;
    return {((nb - 1) << 1) + ((v >> raw_bits) & 1), raw_bits, v & ((1u << raw_bits) - 1)};
}
static inline uint32_t join_distance(uint32_t slot, uint32_t raw_bits, uint32_t raw) {     // inverse, minus the +1
    return ((2 + (slot & 1)) << raw_bits) + raw;
}

struct StreamModel {
    RecentDistances recent;
    Table<2> command;
    Table<4> lit_hi, lit_lo[16];
    Table<3> len_head;                        // length - shortest_len, 7 = escape
    Table<4> len_tail_hi, len_tail_lo[16];    // escaped remainder, two nibbles
    Table<3> slot_hi[4], slot_lo[4][8];       // distance slot, context = min(length - shortest, 3)

    void reset() {
        recent.reset();
        command.reset();
        lit_hi.reset();
        for (auto &t : lit_lo) t.reset();
        len_head.reset();
        len_tail_hi.reset();
        for (auto &t : len_tail_lo) t.reset();
        for (int c = 0; c < 4; c++) {
            slot_hi[c].reset();
            for (auto &t : slot_lo[c]) t.reset();
        }
    }

    uint32_t price_literal(int byte) const {
        int hi = byte >> 4, lo = byte & 15;
        return command.price(kLiteral) + lit_hi.price(hi) + lit_lo[hi].price(lo);
    }
    // length part shared by matches and repeats; excess = len - shortest_len(dist)
    uint32_t price_length(uint32_t excess) const {
        uint32_t p = len_head.price(excess < 7 ? (int)excess : 7);
        if (excess >= 7) {
            uint32_t t = excess - 7;
            p += len_tail_hi.price((int)(t >> 4)) + len_tail_lo[t >> 4].price((int)(t & 15));
        }
        return p;
    }
    uint32_t price_match(uint32_t dist, uint32_t len) const {
        uint32_t excess = len - shortest_len(dist);
        uint32_t ctx = excess < 3 ? excess : 3;
        DistCode dc = split_distance(dist);
        return command.price(kMatch) + price_length(excess) + (dc.raw_bits << kPriceShift) +
               slot_hi[ctx].price((int)(dc.slot >> 3)) + slot_lo[ctx][dc.slot >> 3].price((int)(dc.slot & 7));
    }
    uint32_t price_repeat(uint32_t dist, uint32_t len) const {
        return command.price(kRepeat) + price_length(len - shortest_len(dist)) + (2u << kPriceShift);
    }
};

}  // namespace nlzm_host
#endif
