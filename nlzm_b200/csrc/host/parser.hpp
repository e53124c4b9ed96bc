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

class Staircase {
  public:
    uint32_t top = 0;                                 // longest known length (0 = nothing)
    Staircase() : buf_(kSlide + kLenMax + 2) {}
    uint32_t &operator[](uint32_t len) { return buf_[at_ + len]; }

    // the interface the engine shim drives (MatchTable::Update, NLZM.cpp:848-863)
    void Update(uint32_t dist, uint16_t len) {
        uint32_t *d = &buf_[at_];
        uint32_t known = top < len ? top : len;
        for (uint32_t i = 0; i <= known; i++)
            if (dist < d[i]) d[i] = dist;
        for (uint32_t i = known + 1; i <= len; i++) d[i] = dist;
        if (len > top) top = len;
    }
    // one position forward (MatchTable::CarryFrom with shift 1, NLZM.cpp:836-846)
    void advance() {
        if (top <= 1) { top = 0; return; }
        --top;
        if (++at_ == kSlide) {
            memmove(&buf_[0], &buf_[at_], (top + 1) * sizeof(uint32_t));
            at_ = 0;
        }
    }

  private:
    static constexpr uint32_t kSlide = 1u << 14;
    std::vector<uint32_t> buf_;
    uint32_t at_ = 0;
};

struct ParsedCommand {
    uint8_t kind;        // Command
    uint16_t len;        // 0 for literals
    uint32_t value;      // match: distance; repeat: index into the recent distances
};

// Finders: anything with `template<class T> void FindAndUpdate(T &staircase, uint64_t abs_pos)`
// (GpuMatchFinders of include/nlzm_mf_shim.hpp).
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
        recent_[0] = m.recent;
        node_[1] = {kUnreached, 0, 0, 0, kLiteral};
        recent_[1] = recent_[0];

        uint32_t p = 0, end = 1;
        for (; p < end; ++p) {
            const Node from = node_[p];
            const RecentDistances &from_recent = recent_[p & kRingMask];

            relax(p + 1, from.price + m.price_literal(here[p]), p, kLiteral, 0, 0, from_recent, 0);

            st.advance();
            if (st.top > 0) {
                // re-extend the longest carried candidate against the text
                const uint32_t dist = st[st.top];
                const uint8_t *src = here + p - dist;
                while (st.top < kLenMax && visible > st.top + p && src[st.top] == here[p + st.top]) {
                    ++st.top;
                    st[st.top] = dist;
                }
            }
            if ((st.top < kLongEnough || !(p & kSparseMask)) && visible >= 4 + p)
                finders_.FindAndUpdate(st, start + p);

            uint32_t top = st.top < limit - p ? st.top : limit - p;
            if (top < kLenMin) top = 0;
            while (end < top + p) node_[++end].price = kUnreached;

            uint32_t met = 0;                         // recent distances seen among the candidates
            const uint32_t step = top >= kLenMin + 16 ? (top - kLenMin) >> 4 : 1;
            for (uint32_t len = top; len >= kLenMin; len = len > step ? len - step : 0) {
                const uint32_t dist = st[len];
                if (len < shortest_len(dist)) continue;
                relax(p + len, from.price + m.price_match(dist, len), p, kMatch, len, dist, from_recent, dist);
                int r = from_recent.index_of(dist);
                if (r < 0) continue;
                met |= 1u << r;
                relax(p + len, from.price + m.price_repeat(dist, len), p, kRepeat, len, (uint32_t)r, from_recent, dist);
            }
            if (met != 15) {
                const uint64_t here_shifted = shifted_start + p;
                for (int r = 0; r < 4; r++) {
                    const uint32_t dist = from_recent.d[r];
                    if ((met >> r & 1) || dist >= here_shifted) continue;
                    const uint8_t *src = here + p - dist;
                    uint32_t cap = limit - p, len = 0;
                    while (len < cap && src[len] == here[p + len]) ++len;
                    if (len > kLenMax) len = kLenMax;
                    if (len < shortest_len(dist)) continue;
                    while (end < len + p) node_[++end].price = kUnreached;
                    relax(p + len, from.price + m.price_repeat(dist, len), p, kRepeat, len, (uint32_t)r, from_recent, dist);
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
    void relax(uint32_t to, uint32_t price, uint32_t from, uint8_t kind, uint32_t len, uint32_t value,
               const RecentDistances &from_recent, uint32_t remember) {
        Node &n = node_[to];
        if (!(n.price > price)) return;
        n = {price, (uint16_t)from, (uint16_t)len, value, kind};
        RecentDistances &r = recent_[to & kRingMask];
        r = from_recent;
        if (remember) r.remember(remember);
    }

    const uint8_t *x_;
    Finders &finders_;
    std::vector<Node> node_;
    RecentDistances recent_[kRingMask + 1];
};

}  // namespace nlzm_host
#endif
