// pipeline.hpp — the chunk loop of the encoder: segments are parsed, then coded, one frame per chunk.
//
// Host side of the engine boundary. Restates the driver part of the reference's
// encode_file (NLZM.cpp:1779-1855) and the command coding (model_encode_literal / _match / _rep,
// NLZM.cpp:1274-1367, 1428-1439) for a flat, fully resident input. The reference's ring shifts
// (NLZM.cpp:1786-1792) survive only as the coordinate offset the recent-distance reach test needs.
// Templated on the candidate source so that the same code runs over the engine (codec.cpp) and over
// recorded candidates (tools/host_bench.cpp).
#ifndef NLZM_HOST_PIPELINE_HPP
#define NLZM_HOST_PIPELINE_HPP

#include "frame_coder.hpp"
#include "parser.hpp"
#include "stream_model.hpp"

namespace nlzm_host {

struct BadCandidate {                      // thrown by encode_stream: a (distance, length) that is not a match in the text
    uint64_t pos;
    uint32_t dist, len;
};

struct EncodeCounters {
    uint64_t literals = 0, matches = 0, reps = 0, frames = 0, parses = 0;
};

inline void write_length(FrameWriter &w, StreamModel &m, uint32_t excess) {
    int head = excess < 7 ? (int)excess : 7;
    w.put(m.len_head, head);
    m.len_head.adapt(head);
    if (excess >= 7) {
        int hi = (int)((excess - 7) >> 4), lo = (int)((excess - 7) & 15);
        w.put(m.len_tail_hi, hi);
        w.put(m.len_tail_lo[hi], lo);
        m.len_tail_hi.adapt(hi);
        m.len_tail_lo[hi].adapt(lo);
    }
}

inline void write_command(FrameWriter &w, StreamModel &m, const ParsedCommand &c, uint8_t literal) {
    w.put(m.command, c.kind);
    m.command.adapt(c.kind);
    if (c.kind == kLiteral) {
        int hi = literal >> 4, lo = literal & 15;
        w.put(m.lit_hi, hi);
        w.put(m.lit_lo[hi], lo);
        m.lit_hi.adapt(hi);
        m.lit_lo[hi].adapt(lo);
    } else if (c.kind == kMatch) {
        const uint32_t dist = c.value, excess = c.len - shortest_len(dist), ctx = excess < 3 ? excess : 3;
        write_length(w, m, excess);
        const DistCode dc = split_distance(dist);
        const int hi = (int)(dc.slot >> 3), lo = (int)(dc.slot & 7);
        w.put(m.slot_hi[ctx], hi);
        w.put(m.slot_lo[ctx][hi], lo);
        m.slot_hi[ctx].adapt(hi);
        m.slot_lo[ctx][hi].adapt(lo);
        if (dc.raw_bits > 0) {
            // up to 3 raw bits travel as one field; longer tails as (all but the low nibble), (low nibble)
            if (dc.raw_bits < 4) {
                w.put_raw(dc.raw, dc.raw_bits);
            } else {
                if (dc.raw_bits > 4) w.put_raw(dc.raw >> 4, dc.raw_bits - 4);
                w.put_raw(dc.raw & 15, 4);
            }
        }
        m.recent.remember(dist);
    } else {
        const uint32_t dist = m.recent.d[c.value];
        write_length(w, m, c.len - shortest_len(dist));
        w.put_raw(c.value, 2);
        m.recent.remember(dist);
    }
}

// Appends the frames of in[0, n) to out (the 4-byte stream header and the end marker are the caller's).
template <class Finders>
void encode_stream(const uint8_t *in, uint64_t n, uint32_t hist_bits, uint32_t chunk_size, uint32_t feed_size,
                   Finders &finders, std::vector<uint8_t> &out, EncodeCounters &ec) {
    StreamModel model;
    model.reset();
    FrameWriter frame;
    SegmentParser<Finders> parser(in, finders);
    std::vector<ParsedCommand> cmds;

    for (uint64_t base = 0; base < n; base += chunk_size) {
        const uint64_t coded_end = base + chunk_size < n ? base + chunk_size : n;
        const uint64_t feed_end = base + feed_size < n ? base + feed_size : n;
        // the reference's ring is rebased by one window at a chunk start once it holds two
        const uint64_t windows = base >> hist_bits;
        const uint64_t rebase = (windows > 1 ? windows - 1 : 0) << hist_bits;

        frame.begin();
        for (uint64_t p = base; p < coded_end;) {
            cmds.clear();
            parser.parse(model, p, p - rebase, (uint32_t)(coded_end - p), (uint32_t)(feed_end - p), cmds);
            ++ec.parses;
            for (const ParsedCommand &c : cmds) {
                // every copy the stream is about to promise is compared with the text first (the reference does the
                // same, NLZM.cpp:1824,1839): a wrong candidate from the engine must not become a stream that decodes
                // to different bytes. O(n) in total.
                if (c.kind != kLiteral) {
                    const uint32_t dist = c.kind == kMatch ? c.value : model.recent.d[c.value];
                    if (dist == 0 || dist > p || p + c.len > n || memcmp(in + p - dist, in + p, c.len) != 0)
                        throw BadCandidate{p, dist, c.len};
                }
                write_command(frame, model, c, in[p]);
                if (c.kind == kLiteral) { ++ec.literals; ++p; }
                else { c.kind == kMatch ? ++ec.matches : ++ec.reps; p += c.len; }
            }
        }
        frame.end(out);
        ++ec.frames;
    }
}

}  // namespace nlzm_host
#endif
