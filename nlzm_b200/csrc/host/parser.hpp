// parser.hpp — the forward price-minimising parser that consumes the engine's candidate steps.
//
// Host side of the engine boundary. Restates the reference's parse_table
// (NLZM.cpp:1464-1651) and the per-position candidate table it maintains (MatchTable,
// NLZM.cpp:746-753, 836-863) over a flat, fully resident input instead of ring + lookahead:
//
//   * Staircase: for every length L <= top, the nearest distance known to reach L. Moving one
//     position forward drops the first byte of every candidate (entry L+1 becomes entry L) — here a
//     pointer increment on a sliding buffer instead of the reference's two 1 KB copies per position.
//     The longest carried candidate is then re-extended against the text, and the engine's steps
//     for the position are folded in with min().
//   * Segment: a shortest-path pass over positions [0, end) where `end` grows to the furthest byte
//     any candidate reaches (capped at 4096 and at the chunk's end); it stops where every path has
//     converged. Prices come from the model as it stands at the start of the segment. From every
//     position: a literal; the candidate lengths top, top - step, ... with step = max(1,(top-2)>>4),
//     each as a match and, if its distance is one of the four recent ones, as a repeat; and the
//     recent distances not met that way, measured against the text.
//   * The caller's skip rule: when the carried candidate is already >= 64 long the engine is asked
//     only at every 8th position of the segment (NLZM.cpp:1514,1529).
//
// Ties keep the earlier relaxation (strict '>'), as the reference does; every comparison is on
// uint32 prices with 0xFFFFFFFF as "unreached".
#ifndef NLZM_HOST_PARSER_HPP
#define NLZM_HOST_PARSER_HPP

#include "stream_model.hpp"
#include <string.h>
#include <vector>

namespace nlzm_host {

constexpr uint32_t kSegmentMax = 1u << 12;
constexpr uint32_t kLongEnough = 64;         // carried candidate this long: ask the engine sparsely
constexpr uint32_t kSparseMask = 7;

// What the parser needs from a candidate's distance before it can price it, packed next to the distance so that
// both move through the staircase together: bits 0..5 distance slot, bits 6..7 shortest length - 2. The engine
// computes it on the GPU for every step it emits (nlzm_mf_step); candidates that come from the host
// (one Update per step through the reference-shaped interface) get it from split_distance / shortest_len.
static inline uint32_t candidate_code(uint32_t dist) {
    return split_distance(dist).slot | ((shortest_len(dist) - kLenMin) << 6);
}
static inline uint64_t candidate_entry(uint32_t dist, uint32_t code) { return ((uint64_t)dist << 32) | code; }

class Staircase {
  public:
    uint32_t top = 0;                                 // longest known length (0 = nothing)
    Staircase() : buf_(kSlide + kLenMax + 2) {}
    // entry of a length: distance in the high word (entries compare like distances), its code in the low word
    uint64_t entry(uint32_t len) const { return buf_[at_ + len]; }
    uint32_t operator[](uint32_t len) const { return (uint32_t)(buf_[at_ + len] >> 32); }
    void set(uint32_t len, uint64_t e) { buf_[at_ + len] = e; }

    // the interface the engine shim drives (MatchTable::Update, NLZM.cpp:848-863)
    void Update(uint32_t dist, uint16_t len) {
        const uint64_t e = candidate_entry(dist, candidate_code(dist));
        uint64_t *d = &buf_[at_];
        uint32_t known = top < len ? top : len;
        for (uint32_t i = 0; i <= known; i++)
            if (e < d[i]) d[i] = e;
        for (uint32_t i = known + 1; i <= len; i++) d[i] = e;
        if (len > top) top = len;
    }
    // A whole position's steps in one pass: the same result as Update(dist_j, len_j) for j = 0..n-1
    // when the steps are strictly increasing in len and dist (the engine's contract, nlzm_mf.h) —
    // step j then only matters for lengths above len_{j-1}, so the work is O(longest) instead of
    // O(sum of lengths). entry_of(j) = candidate_entry(distance, code) of step j.
    template <class EntryOf, class LenOf> void merge_steps(uint32_t n, EntryOf entry_of, LenOf len_of) {
        if (n == 0) return;
        uint64_t *d = &buf_[at_];
        uint32_t i = 0, len = 0;
        for (uint32_t j = 0; j < n; j++) {
            const uint64_t e = entry_of(j);
            len = len_of(j);
            const uint32_t known = top < len ? top : len;
            for (; i <= known; i++) d[i] = e < d[i] ? e : d[i];
            for (; i <= len; i++) d[i] = e;
        }
        if (len > top) top = len;
    }
    // one position forward (MatchTable::CarryFrom with shift 1, NLZM.cpp:836-846)
    void advance() {
        if (top <= 1) { top = 0; return; }
        --top;
        if (++at_ == kSlide) {
            memmove(&buf_[0], &buf_[at_], (top + 1) * sizeof(uint64_t));
            at_ = 0;
        }
    }

  private:
    static constexpr uint32_t kSlide = 1u << 14;
    std::vector<uint64_t> buf_;
    uint32_t at_ = 0;
};

struct ParsedCommand {
    uint8_t kind;        // Command
    uint16_t len;        // 0 for literals
    uint32_t value;      // match: distance; repeat: index into the recent distances
};

// Finders: anything with `template<class T> void FindAndUpdate(T &staircase, uint64_t abs_pos)` that
// folds the candidates of abs_pos into the staircase — BlockFeed (codec.cpp) over the engine, or
// GpuMatchFinders of include/nlzm_mf_shim.hpp (which calls Staircase::Update once per step).
template <class Finders> class SegmentParser {
  public:
    SegmentParser(const uint8_t *text, Finders &finders) : x_(text), finders_(finders), node_(kSegmentMax + 1) {}

    Staircase carried;                                // survives across segments and chunks (mt_carry)

    // Parses a segment starting at absolute offset `start`, which is `shifted_start` in the
    // reference's ring coordinates (only the recent-distance reach test reads it). `limit` = bytes
    // left in the chunk's coded range, `visible` = bytes of text the encoder may look at from `start`
    // (the chunk's feed). Appends the commands to `out` in stream order; returns the bytes they cover.
    uint32_t parse(const StreamModel &m, uint64_t start, uint64_t shifted_start, uint32_t limit, uint32_t visible,
                   std::vector<ParsedCommand> &out) {
        if (limit > kSegmentMax) limit = kSegmentMax;
        const uint8_t *here = x_ + start;
        Staircase &st = carried;

        node_[0] = {0, 0, 0, 0, kNone};
        node_[1] = {kUnreached, 0, 0, 0, kLiteral};

        ++stamp_;                                     // prices are fixed for the segment: memoise them
        const uint32_t match_cmd = m.command.price(kMatch), repeat_cmd = m.command.price(kRepeat) + (2u << kPriceShift);

        uint32_t p = 0, end = 1;
        for (; p < end; ++p) {
            const Node from = node_[p];
            // the recent distances on the best path to p: those of its predecessor, plus the distance
            // of the match that leads here (repeats and literals change nothing). Derived when p is
            // visited — its predecessor is final by then — rather than copied at every relaxation.
            RecentDistances &from_recent = recent_[p & kRingMask];
            if (p == 0) {
                from_recent = m.recent;
            } else {
                from_recent = recent_[from.from & kRingMask];
                if (from.kind == kMatch) from_recent.remember(from.value);
            }

            relax(p + 1, from.price + literal_price(m, here[p]), p, kLiteral, 0, 0);

            st.advance();
            if (st.top > 0) {
                // re-extend the longest carried candidate against the text
                const uint64_t longest = st.entry(st.top);
                const uint8_t *src = here + p - (uint32_t)(longest >> 32);
                while (st.top < kLenMax && visible > st.top + p && src[st.top] == here[p + st.top]) {
                    ++st.top;
                    st.set(st.top, longest);
                }
            }
            if ((st.top < kLongEnough || !(p & kSparseMask)) && visible >= 4 + p)
                finders_.FindAndUpdate(st, start + p);

            // the next position re-extends the longest candidate: one byte past its end, in text this
            // thread may never have touched (the GPU found the match) — fetch it a position ahead
            if (st.top > 0) __builtin_prefetch(here + p - st[st.top] + st.top);

            uint32_t top = st.top < limit - p ? st.top : limit - p;
            if (top < kLenMin) top = 0;
            while (end < top + p) node_[++end].price = kUnreached;

            uint32_t met = 0;                         // recent distances seen among the candidates
            const uint32_t step = top >= kLenMin + 16 ? (top - kLenMin) >> 4 : 1;
            uint64_t run = 0;                                                 // per run of equal distances
            uint32_t dist = 0, shortest = 0, slot = 0, raw_price = 0;
            int recent_index = -1;
            for (uint32_t len = top; len >= kLenMin; len = len > step ? len - step : 0) {
                const uint64_t e = st.entry(len);
                if (e != run) {
                    // slot and shortest length come with the candidate (computed on the GPU)
                    run = e;
                    dist = (uint32_t)(e >> 32);
                    slot = (uint32_t)e & 63u;
                    shortest = kLenMin + (((uint32_t)e >> 6) & 3u);
                    raw_price = (slot < 4 ? 0u : (slot >> 1) - 1u) << kPriceShift;
                    recent_index = from_recent.index_of(dist);
                }
                if (len < shortest) continue;
                const uint32_t excess = len - shortest;
                const uint32_t length_part = length_price(m, excess);
                relax(p + len, from.price + match_cmd + length_part + raw_price + slot_price(m, excess < 3 ? excess : 3, slot),
                      p, kMatch, len, dist);
                if (recent_index < 0) continue;
                met |= 1u << recent_index;
                relax(p + len, from.price + repeat_cmd + length_part, p, kRepeat, len, (uint32_t)recent_index);
            }
            if (met != 15 && limit - p >= kLenMin) {
                const uint64_t here_shifted = shifted_start + p;
                const uint32_t cap = limit - p < kLenMax ? limit - p : kLenMax;     // longer is clamped anyway
                const uint16_t head = load16(here + p);
                for (int r = 0; r < 4; r++) {
                    const uint32_t dist = from_recent.d[r];
                    if ((met >> r & 1) || dist >= here_shifted) continue;
                    const uint8_t *src = here + p - dist;
                    if (load16(src) != head) continue;                  // every usable length is >= 2
                    const uint32_t len = 2 + common_prefix(src + 2, here + p + 2, cap - 2);
                    if (len < shortest_len(dist)) continue;
                    while (end < len + p) node_[++end].price = kUnreached;
                    relax(p + len, from.price + repeat_cmd + length_price(m, len - shortest_len(dist)), p, kRepeat, len, (uint32_t)r);
                }
            }
        }

        // walk the chosen path back from the end, then emit it forwards
        const size_t at = out.size();
        for (uint32_t cur = end; cur != 0; cur = node_[cur].from)
            out.push_back({node_[cur].kind, node_[cur].len, node_[cur].value});
        for (size_t i = at, j = out.size(); i + 1 < j; ++i, --j) {
            ParsedCommand t = out[i]; out[i] = out[j - 1]; out[j - 1] = t;
        }
        return end;
    }

  private:
    static constexpr uint32_t kUnreached = 0xFFFFFFFFu;
    static constexpr uint32_t kRingMask = 0x1FF;      // recent-distance states live in a ring of 512 positions
    struct Node {
        uint32_t price;
        uint16_t from, len;
        uint32_t value;
        uint8_t kind;
    };
    void relax(uint32_t to, uint32_t price, uint32_t from, uint8_t kind, uint32_t len, uint32_t value) {
        Node &n = node_[to];
        if (n.price > price) n = {price, (uint16_t)from, (uint16_t)len, value, kind};
    }

    static uint16_t load16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
    static uint32_t common_prefix(const uint8_t *a, const uint8_t *b, uint32_t cap) {
        uint32_t n = 0;
        while (n + 8 <= cap) {
            uint64_t va, vb;
            memcpy(&va, a + n, 8);
            memcpy(&vb, b + n, 8);
            if (va != vb) return n + (uint32_t)(__builtin_ctzll(va ^ vb) >> 3);
            n += 8;
        }
        while (n < cap && a[n] == b[n]) ++n;
        return n;
    }
    // memoised prices, valid while stamp matches the current segment
    struct Memo { uint32_t stamp, price; };
    uint32_t literal_price(const StreamModel &m, uint8_t y) {
        Memo &c = lit_memo_[y];
        if (c.stamp != stamp_) c = {stamp_, m.price_literal(y)};
        return c.price;
    }
    uint32_t length_price(const StreamModel &m, uint32_t excess) {
        Memo &c = len_memo_[excess];
        if (c.stamp != stamp_) c = {stamp_, m.price_length(excess)};
        return c.price;
    }
    uint32_t slot_price(const StreamModel &m, uint32_t ctx, uint32_t slot) {
        Memo &c = slot_memo_[ctx][slot];
        if (c.stamp != stamp_) c = {stamp_, m.slot_hi[ctx].price((int)(slot >> 3)) + m.slot_lo[ctx][slot >> 3].price((int)(slot & 7))};
        return c.price;
    }

    const uint8_t *x_;
    Finders &finders_;
    std::vector<Node> node_;
    RecentDistances recent_[kRingMask + 1];
    uint32_t stamp_ = 0;
    Memo lit_memo_[256] = {}, len_memo_[kLenMax + 1] = {}, slot_memo_[4][64] = {};
};

}  // namespace nlzm_host
#endif
