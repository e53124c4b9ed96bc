// codec.cpp — libnlzm_codec: the host pipeline around the B200 match-finding engine.
//
// SURVEY.md §8 f1 (parser consumer + pipeline), f2 (stream writer), f4 (stream reader). C ABI in
// include/nlzm_codec.h. Restates the reference's encode_file / decode_file drivers
// (NLZM.cpp:1711-1910, 1912-2039) for a flat input: the whole file goes to HBM once
// (nlzm_mf_set_input), candidate blocks come back double-buffered (the GPU computes block N+1
// while this thread parses and codes block N), and per chunk of `chunk_size` bytes one frame is
// written. Ring shifts survive only as the coordinate offset the recent-distance reach test needs.
//
// The matcher stage has no CPU implementation here: without a CUDA device nlzm_codec_compress
// fails with NLZM_CODEC_E_ENGINE.
#include "../../../include/nlzm_codec.h"
#include "../../../include/nlzm_mf.h"
#include "frame_coder.hpp"
#include "parser.hpp"
#include "pipeline.hpp"
#include "stream_model.hpp"

#include <chrono>
#include <future>
#include <vector>
#include <new>
#include <stdlib.h>
#include <string>

using namespace nlzm_host;

namespace {

thread_local std::string g_error;

int fail(int rc, const std::string &why) {
    g_error = why;
    return rc;
}

double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

struct EngineError {
    std::string what;
};

// The candidate source: the same double buffering as GpuMatchFinders (include/nlzm_mf_shim.hpp) —
// while this thread parses and codes block N out of slot N&1, the engine already computes block N+1
// into the other slot — written against the C ABI directly so that a failing engine call is
// reported to the caller (the shim keeps the reference's ASSERT -> exit(-1) behaviour). A position's
// steps are folded into the staircase in one pass (Staircase::merge_steps), and the time spent
// blocked on a hand-over is accounted for.
class BlockFeed {
  public:
    ~BlockFeed() { close(); }
    uint64_t steps_served = 0, blocks_fetched = 0;
    double ms_wait = 0;

    void open(const nlzm_mf_config &mc, const uint8_t *in) {
        int rc = nlzm_mf_create(&mc, &mf_);
        if (rc) throw EngineError{std::string("engine create failed: ") + nlzm_mf_last_error(nullptr)};
        check(nlzm_mf_set_input(mf_, in, mc.file_len), "set_input");
        flen_ = mc.file_len;
        block_ = mc.max_range;
        submit_next();
    }
    void close() {
        if (!mf_) return;
        if (pending_) { nlzm_mf_view v; nlzm_mf_fetch(mf_, slot_ ^ 1, &v); }
        nlzm_mf_destroy(mf_);
        mf_ = nullptr;
    }
    template <class T> void FindAndUpdate(T &st, uint64_t abs_pos) {
        if (abs_pos >= view_.end) {
            auto t0 = std::chrono::steady_clock::now();
            while (abs_pos >= view_.end) advance();
            ms_wait += ms_since(t0);
        }
        const uint64_t i = abs_pos - view_.begin;
        const nlzm_mf_step *s = view_.steps + view_.offsets[i];
        const uint32_t n = view_.offsets[i + 1] - view_.offsets[i];
        // distance + the engine's pre-pricing (slot, shortest length) as one staircase entry
        st.merge_steps(n,
                       [s](uint32_t j) {
                           return nlzm_host::candidate_entry(NLZM_MF_STEP_DIST(s[j]), NLZM_MF_STEP_SLOT(s[j]) | ((NLZM_MF_STEP_SHORTEST(s[j]) - 2u) << 6));
                       },
                       [s](uint32_t j) { return NLZM_MF_STEP_LEN(s[j]); });
        steps_served += n;
    }

  private:
    static constexpr uint64_t kFirstBlockMin = 64u << 10;   // default block 32 Mi -> first block 4 Mi
    void check(int rc, const char *what) {
        if (rc) throw EngineError{std::string("engine ") + what + " failed (rc " + std::to_string(rc) + "): " + nlzm_mf_last_error(mf_)};
    }
    void submit_next() {
        if (next_begin_ >= flen_) return;
        // the first block is short so that parsing starts almost at once; it has no window behind it
        // to rebuild, so the cut costs the engine nothing
        uint64_t first = block_ / 8 > kFirstBlockMin ? block_ / 8 : kFirstBlockMin;
        const uint64_t want = next_begin_ == 0 && first < block_ ? first : block_;
        const uint64_t e = next_begin_ + want < flen_ ? next_begin_ + want : flen_;
        check(nlzm_mf_submit(mf_, next_begin_, e, slot_ ^ 1), "submit");
        next_begin_ = e;
        pending_ = true;
    }
    void advance() {
        if (!pending_) throw EngineError{"position past the end of the input"};
        slot_ ^= 1;
        pending_ = false;
        check(nlzm_mf_fetch(mf_, slot_, &view_), "fetch");
        ++blocks_fetched;
        submit_next();
    }
    nlzm_mf *mf_ = nullptr;
    nlzm_mf_view view_{};
    uint64_t flen_ = 0, block_ = 0, next_begin_ = 0;
    int slot_ = 1;
    bool pending_ = false;
};

// The same candidate source over SEVERAL engines of one process (one per GPU): the input is replicated, the
// blocks of the file go round-robin to the engines, and the parser still sees one ascending stream of positions
// (parse_table's loop, NLZM.cpp:1486-1543). A batch of N blocks is in flight at a time:
//   1. every engine ranks and merges its block alone (nlzm_mf_prepare, N host threads, N GPUs busy),
//   2. block j takes the sorted blocks of the window behind it from the engine that owns block j-1
//      (nlzm_mf_export_segments / nlzm_mf_import_segment: a peer copy over NVLink inside one process),
//   3. every engine finishes its block (nlzm_mf_submit) and copies the records to its pinned buffers,
// while the parser consumes the previous batch. Blocks are at least one window long, so only block j-1 lies
// behind block j. Result slot = parity of the batch.
class MultiBlockFeed {
  public:
    ~MultiBlockFeed() { close(); }
    uint64_t steps_served = 0, blocks_fetched = 0;
    double ms_wait = 0;

    void open(nlzm_mf_config mc, const int32_t *devices, uint32_t n_dev, uint32_t window, const uint8_t *in) {
        flen_ = mc.file_len;
        window_ = window;
        uint64_t block = mc.max_range < window ? window : mc.max_range;
        if (block > (1ull << 28)) block = 1ull << 28;
        mc.max_range = block;
        // block 0 is short (parsing starts almost at once); every later block is a full one
        uint64_t first = block / 8 > (64u << 10) ? block / 8 : (64u << 10);
        for (uint64_t b = 0; b < flen_;) {
            const uint64_t want = b == 0 && first < block ? first : block;
            const uint64_t e = b + want < flen_ ? b + want : flen_;
            blocks_.push_back({b, e});
            b = e;
        }
        for (uint32_t d = 0; d < n_dev; d++) {
            mc.device = devices[d];
            nlzm_mf *h = nullptr;
            int rc = nlzm_mf_create(&mc, &h);
            if (rc) throw EngineError{std::string("engine create failed on device ") + std::to_string(devices[d]) + ": " + nlzm_mf_last_error(nullptr)};
            mf_.push_back(h);
            check(h, nlzm_mf_set_input(h, in, mc.file_len), "set_input");
        }
        launch_batch(0);
    }
    void close() {
        if (launching_.valid()) { try { launching_.get(); } catch (...) {} }
        for (size_t j = 0; j < submitted_.size(); j++)
            if (submitted_[j] && !fetched_[j]) { nlzm_mf_view v; nlzm_mf_fetch(mf_[j % mf_.size()], slot_of(j), &v); }
        for (nlzm_mf *h : mf_) nlzm_mf_destroy(h);
        mf_.clear();
    }
    template <class T> void FindAndUpdate(T &st, uint64_t abs_pos) {
        if (abs_pos >= view_.end) {
            auto t0 = std::chrono::steady_clock::now();
            while (abs_pos >= view_.end) advance();
            ms_wait += ms_since(t0);
        }
        const uint64_t i = abs_pos - view_.begin;
        const nlzm_mf_step *s = view_.steps + view_.offsets[i];
        const uint32_t n = view_.offsets[i + 1] - view_.offsets[i];
        st.merge_steps(n,
                       [s](uint32_t j) {
                           return nlzm_host::candidate_entry(NLZM_MF_STEP_DIST(s[j]), NLZM_MF_STEP_SLOT(s[j]) | ((NLZM_MF_STEP_SHORTEST(s[j]) - 2u) << 6));
                       },
                       [s](uint32_t j) { return NLZM_MF_STEP_LEN(s[j]); });
        steps_served += n;
    }

  private:
    struct Range { uint64_t b, e; };
    static void check(nlzm_mf *h, int rc, const char *what) {
        if (rc) throw EngineError{std::string("engine ") + what + " failed (rc " + std::to_string(rc) + "): " + nlzm_mf_last_error(h)};
    }
    int slot_of(size_t j) const { return (int)((j / mf_.size()) & 1); }

    // steps 1-3 for blocks [k*N, k*N+N), on a helper thread so that the parser keeps going
    void launch_batch(size_t k) {
        const size_t n_dev = mf_.size(), j0 = k * n_dev;
        if (j0 >= blocks_.size()) return;
        submitted_.resize(blocks_.size(), false);
        fetched_.resize(blocks_.size(), false);
        launching_ = std::async(std::launch::async, [this, j0, n_dev]() {
            const size_t j1 = j0 + n_dev < blocks_.size() ? j0 + n_dev : blocks_.size();
            std::vector<std::future<int>> prep;
            for (size_t j = j0; j < j1; j++)
                prep.push_back(std::async(std::launch::async, [this, j]() { return nlzm_mf_prepare(mf_[j % mf_.size()], blocks_[j].b, blocks_[j].e); }));
            for (size_t j = j0; j < j1; j++) check(mf_[j % n_dev], prep[j - j0].get(), "prepare");
            for (size_t j = j0 > 0 ? j0 : 1; j < j1; j++) {
                nlzm_mf *from = mf_[(j - 1) % n_dev], *to = mf_[j % n_dev];
                if (from == to) continue;                       // one engine: its own retained segments
                uint32_t n_seg = 0;
                check(from, nlzm_mf_export_segments(from, nullptr, 0, &n_seg), "export_segments");
                std::vector<nlzm_mf_segment> segs(n_seg ? n_seg : 1);
                check(from, nlzm_mf_export_segments(from, segs.data(), n_seg, &n_seg), "export_segments");
                const uint64_t need = blocks_[j].b > (uint64_t)(window_ - 1) ? blocks_[j].b - (window_ - 1) : 0;
                for (uint32_t i = 0; i < n_seg; i++)
                    if (segs[i].pos_end > need && segs[i].pos_end <= blocks_[j].b && segs[i].pos_begin >= blocks_[j - 1].b)
                        check(to, nlzm_mf_import_segment(to, &segs[i], 0), "import_segment");
            }
            for (size_t j = j0; j < j1; j++) {
                check(mf_[j % n_dev], nlzm_mf_submit(mf_[j % n_dev], blocks_[j].b, blocks_[j].e, slot_of(j)), "submit");
                submitted_[j] = true;
            }
        });
    }
    void advance() {
        if (next_ >= blocks_.size()) throw EngineError{"position past the end of the input"};
        const size_t n_dev = mf_.size();
        if (next_ % n_dev == 0) {
            // first block of a batch: its launch must be complete; the batch after it may start now, because
            // the slots it will write (same parity as the batch before this one) have been consumed
            if (launching_.valid()) launching_.get();
            launch_batch(next_ / n_dev + 1);
        }
        if (!submitted_[next_]) {                               // launched by the call above (next batch) — wait for it
            if (launching_.valid()) launching_.get();
        }
        check(mf_[next_ % n_dev], nlzm_mf_fetch(mf_[next_ % n_dev], slot_of(next_), &view_), "fetch");
        fetched_[next_] = true;
        ++next_;
        ++blocks_fetched;
    }

    std::vector<nlzm_mf *> mf_;
    std::vector<Range> blocks_;
    std::vector<char> submitted_, fetched_;
    std::future<void> launching_;
    nlzm_mf_view view_{};
    uint64_t flen_ = 0;
    uint32_t window_ = 0;
    size_t next_ = 0;
};

uint32_t read_length_excess(FrameReader &r, StreamModel &m) {
    uint32_t excess = (uint32_t)r.get(m.len_head);
    m.len_head.adapt((int)excess);
    if (excess == 7) {
        int hi = r.get(m.len_tail_hi), lo = r.get(m.len_tail_lo[hi]);
        m.len_tail_hi.adapt(hi);
        m.len_tail_lo[hi].adapt(lo);
        excess += ((uint32_t)hi << 4) + (uint32_t)lo;
    }
    return excess;
}

int hand_over(std::vector<uint8_t> &v, uint8_t **out, uint64_t *out_len) {
    uint8_t *p = (uint8_t *)malloc(v.size() ? v.size() : 1);
    if (!p) return fail(NLZM_CODEC_E_NOMEM, "out of host memory");
    if (!v.empty()) memcpy(p, v.data(), v.size());
    *out = p;
    *out_len = v.size();
    return 0;
}

int compress_impl(const uint8_t *in, uint64_t n, const nlzm_codec_config &cfg, std::vector<uint8_t> &out,
                  nlzm_codec_stats &st) {
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t window_bits = cfg.window_bits < 15 ? 15 : (cfg.window_bits > 28 ? 28 : cfg.window_bits);
    nlzm_mf_geometry g;
    if (nlzm_mf_get_geometry(n, window_bits, &g)) return fail(NLZM_CODEC_E_ARG, "bad geometry");

    out.clear();
    out.reserve((size_t)(n / 2 + 64));
    out.push_back((uint8_t)(g.hist_bits >> 8));
    out.push_back((uint8_t)g.hist_bits);
    out.push_back((uint8_t)(g.frame_bits >> 8));
    out.push_back((uint8_t)g.frame_bits);

    if (n > 0) {
        nlzm_mf_config mc{};
        mc.struct_size = sizeof mc;
        mc.hist_bits = window_bits;
        mc.file_len = n;
        mc.device = cfg.device;
        mc.finder_mask = NLZM_MF_ALL;
        mc.max_range = cfg.block_len ? cfg.block_len : (g.window > (32u << 20) ? g.window : (32u << 20));
        if (mc.max_range > (1ull << 28)) mc.max_range = 1ull << 28;
        const bool has_list = cfg.struct_size >= sizeof(nlzm_codec_config) && cfg.n_devices > 1;
        EncodeCounters ec;
        auto run = [&](auto &finders) {
            encode_stream(in, n, g.hist_bits, g.chunk_size, g.feed_size, finders, out, ec);
            st.steps_served = finders.steps_served;
            st.engine_blocks = finders.blocks_fetched;
            st.ms_engine_wait = finders.ms_wait;
            finders.close();
        };
        if (has_list) {
            if (cfg.n_devices > 8) return fail(NLZM_CODEC_E_ARG, "at most 8 devices");
            MultiBlockFeed finders;
            finders.open(mc, cfg.devices, cfg.n_devices, g.window, in);
            run(finders);
        } else {
            BlockFeed finders;
            finders.open(mc, in);
            run(finders);
        }
        st.literals = ec.literals; st.matches = ec.matches; st.reps = ec.reps;
        st.frames = ec.frames; st.parses = ec.parses;
    }
    out.insert(out.end(), 4, 0);             // a frame with zero ops ends the stream
    st.in_bytes = n;
    st.out_bytes = out.size();
    st.ms_total = ms_since(t0);
    return 0;
}

// Decoder output: raw storage with `size` bytes valid and always >= kSlack bytes of room behind
// them, so that a literal is one store and a match copies in 8-byte strides without further checks.
struct OutputSink {
    static constexpr size_t kSlack = kLenMax + 16;      // longest decodable copy (262 + 5) + one stride
    uint8_t *p = nullptr;
    size_t size = 0, cap = 0;
    ~OutputSink() { free(p); }
    bool room() {
        if (cap - size >= kSlack) return true;
        size_t want = cap + cap / 2 + kSlack + (1u << 16);
        uint8_t *q = (uint8_t *)realloc(p, want);
        if (!q) return false;
        p = q;
        cap = want;
        return true;
    }
    void literal(uint8_t y) { p[size++] = y; }
    void copy(size_t dist, size_t len) {
        uint8_t *to = p + size;
        const uint8_t *from = to - dist;
        if (dist >= 8) {
            for (size_t i = 0; i < len; i += 8) memcpy(to + i, from + i, 8);      // may run up to 7 bytes into the slack
        } else {
            for (size_t i = 0; i < len; i++) to[i] = from[i];                      // overlapping: byte by byte
        }
        size += len;
    }
    uint8_t *release() { uint8_t *r = p; p = nullptr; return r; }
};

int decompress_impl(const uint8_t *in, uint64_t n, OutputSink &out) {
    if (n < 8) return fail(NLZM_CODEC_E_STREAM, "stream shorter than header + end marker");
    const uint32_t hist_bits = ((uint32_t)in[0] << 8) | in[1], frame_bits = ((uint32_t)in[2] << 8) | in[3];
    if (hist_bits < 10 || hist_bits > 28 || frame_bits < 12 || frame_bits > 20)
        return fail(NLZM_CODEC_E_STREAM, "bad stream header");
    StreamModel model;
    model.reset();
    FrameReader frame;
    uint64_t at = 4;
    for (;;) {
        int64_t size = frame.begin(in + at, (size_t)(n - at));
        if (size < 0) return fail(NLZM_CODEC_E_STREAM, "malformed frame header at offset " + std::to_string(at));
        if (size == 0) break;
        while (frame.ops_left() > 0) {
            if (!out.room()) return fail(NLZM_CODEC_E_NOMEM, "out of host memory");
            int kind = frame.get(model.command);
            model.command.adapt(kind);
            if (kind == kLiteral) {
                int hi = frame.get(model.lit_hi), lo = frame.get(model.lit_lo[hi]);
                model.lit_hi.adapt(hi);
                model.lit_lo[hi].adapt(lo);
                out.literal((uint8_t)((hi << 4) | lo));
            } else if (kind == kMatch || kind == kRepeat) {
                uint32_t dist, len;
                if (kind == kMatch) {
                    const uint32_t excess = read_length_excess(frame, model), ctx = excess < 3 ? excess : 3;
                    const int hi = frame.get(model.slot_hi[ctx]), lo = frame.get(model.slot_lo[ctx][hi]);
                    model.slot_hi[ctx].adapt(hi);
                    model.slot_lo[ctx][hi].adapt(lo);
                    uint32_t v = ((uint32_t)hi << 3) | (uint32_t)lo;
                    if (v >= 4) {
                        const uint32_t raw_bits = (v >> 1) - 1;
                        if (raw_bits > 26) return fail(NLZM_CODEC_E_STREAM, "distance slot out of range");
                        uint32_t raw;
                        if (raw_bits < 4) raw = frame.get_raw(raw_bits);
                        else {
                            raw = raw_bits > 4 ? frame.get_raw(raw_bits - 4) << 4 : 0;
                            raw += frame.get_raw(4);
                        }
                        v = join_distance(v, raw_bits, raw);
                    }
                    dist = v + 1;
                    len = excess + shortest_len(dist);
                } else {
                    dist = model.recent.d[frame.get_raw(2)];
                    len = read_length_excess(frame, model) + shortest_len(dist);
                }
                model.recent.remember(dist);
                if (frame.bad()) break;
                if (dist > out.size) return fail(NLZM_CODEC_E_STREAM, "distance reaches before the start of the output");
                out.copy(dist, len);
            } else {
                return fail(NLZM_CODEC_E_STREAM, "unknown command");
            }
            if (frame.bad()) break;
        }
        if (frame.bad()) return fail(NLZM_CODEC_E_STREAM, "frame at offset " + std::to_string(at) + " is truncated");
        at += (uint64_t)size;
        if (n - at < 4) return fail(NLZM_CODEC_E_STREAM, "missing end marker");
    }
    return 0;
}

}  // namespace

extern "C" {

int nlzm_codec_abi_version(void) { return NLZM_CODEC_ABI_VERSION; }

int nlzm_codec_compress(const uint8_t *in, uint64_t in_len, const nlzm_codec_config *cfg, uint8_t **out,
                        uint64_t *out_len, nlzm_codec_stats *stats) {
    if (!cfg || cfg->struct_size != sizeof(nlzm_codec_config) || !out || !out_len || (!in && in_len))
        return fail(NLZM_CODEC_E_ARG, "bad argument");
    if (in_len >= (1ull << 31)) return fail(NLZM_CODEC_E_ARG, "input must be smaller than 2^31 bytes");
    *out = nullptr;
    *out_len = 0;
    nlzm_codec_stats st{};
    try {
        std::vector<uint8_t> v;
        int rc = compress_impl(in, in_len, *cfg, v, st);
        if (rc) return rc;
        rc = hand_over(v, out, out_len);
        if (rc) return rc;
    } catch (const BadCandidate &b) {
        return fail(NLZM_CODEC_E_ENGINE, "candidate (distance " + std::to_string(b.dist) + ", length " + std::to_string(b.len) +
                                             ") at position " + std::to_string(b.pos) + " is not a match in the text");
    } catch (const EngineError &e) {
        return fail(NLZM_CODEC_E_ENGINE, e.what);
    } catch (const std::bad_alloc &) {
        return fail(NLZM_CODEC_E_NOMEM, "out of host memory");
    }
    if (stats) *stats = st;
    return 0;
}

int nlzm_codec_decompress(const uint8_t *in, uint64_t in_len, uint8_t **out, uint64_t *out_len) {
    if (!in || !out || !out_len) return fail(NLZM_CODEC_E_ARG, "bad argument");
    *out = nullptr;
    *out_len = 0;
    try {
        OutputSink sink;
        int rc = decompress_impl(in, in_len, sink);
        if (rc) return rc;
        if (!sink.p && !sink.room()) return fail(NLZM_CODEC_E_NOMEM, "out of host memory");   // empty output still owns a block
        *out_len = sink.size;
        *out = sink.release();
        return 0;
    } catch (const std::bad_alloc &) {
        return fail(NLZM_CODEC_E_NOMEM, "out of host memory");
    }
}

void nlzm_codec_free(uint8_t *p) { free(p); }

const char *nlzm_codec_last_error(void) { return g_error.c_str(); }

}  // extern "C"
