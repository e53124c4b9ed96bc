// frame_coder.hpp — one NLZM frame: a raw bit string plus four interleaved 32-bit rANS states.
//
// Host side of the engine boundary. Restates the reference's frame layout
//   [ops BE32][bit-section bytes incl. this 12-byte header BE32][rANS bytes BE32][bits][rANS]
// and its coder (NLZM.cpp:440-490 rANS step, 492-640 CodeFrame, 642-731 DecodeFrame):
//   * entropy-coded symbols are queued as (low, freq) pairs and coded back to front at the end of
//     the frame, symbol i on state i & 3, 16-bit renormalisation, states start at 2^16;
//   * the rANS section is laid out for a forward-reading decoder: states 0..3 little endian, then
//     the renormalisation words big endian in the order the decoder consumes them;
//   * raw bits are packed most significant bit first; closing the section always appends four
//     bytes (the partial byte, then zeros) because the reader prefetches up to 24 bits.
// `ops` counts every queued symbol and every raw-bit field; the reader stops when it reaches zero.
#ifndef NLZM_HOST_FRAME_CODER_HPP
#define NLZM_HOST_FRAME_CODER_HPP

#include "stream_model.hpp"
#include <algorithm>
#include <vector>

namespace nlzm_host {

constexpr uint32_t kRansFloor = 1u << 16;

class FrameWriter {
  public:
    void begin() {
        queue_.clear();
        bits_.clear();
        acc_ = 0;
        acc_bits_ = 0;
        ops_ = 0;
    }
    template <int BITS> void put(const Table<BITS> &t, int y) {
        queue_.push_back((t.freq(y) << 16) | t.low(y));
        ++ops_;
    }
    void put_raw(uint32_t v, uint32_t nb) {
        acc_ = (acc_ << nb) | v;
        acc_bits_ += nb;
        while (acc_bits_ >= 8) {
            acc_bits_ -= 8;
            bits_.push_back((uint8_t)(acc_ >> acc_bits_));
        }
        ++ops_;
    }
    bool empty() const { return ops_ == 0; }

    // appends the finished frame to out, returns its size
    size_t end(std::vector<uint8_t> &out) {
        bits_.push_back(acc_bits_ ? (uint8_t)(acc_ << (8 - acc_bits_)) : 0);
        bits_.insert(bits_.end(), 3, 0);

        // back to front; tail_ collects bytes in reverse of their final order
        tail_.clear();
        uint32_t x[4] = {kRansFloor, kRansFloor, kRansFloor, kRansFloor};
        for (size_t i = queue_.size(); i-- > 0;) {
            uint32_t freq = queue_[i] >> 16, low = queue_[i] & 0xFFFF;
            uint32_t &s = x[i & 3];
            if (s >= (freq << (32 - kProbBits))) {
                tail_.push_back((uint8_t)s);
                tail_.push_back((uint8_t)(s >> 8));
                s >>= 16;
            }
            s = ((s / freq) << kProbBits) + (s % freq) + low;
        }
        for (int k = 3; k >= 0; k--)
            for (int b = 3; b >= 0; b--) tail_.push_back((uint8_t)(x[k] >> (8 * b)));

        const uint32_t bit_section = 12 + (uint32_t)bits_.size(), rans_section = (uint32_t)tail_.size();
        const size_t at = out.size();
        out.resize(at + bit_section + rans_section);
        uint8_t *p = out.data() + at;
        store_be32(p, ops_);
        store_be32(p + 4, bit_section);
        store_be32(p + 8, rans_section);
        std::copy(bits_.begin(), bits_.end(), p + 12);
        std::reverse_copy(tail_.begin(), tail_.end(), p + bit_section);
        begin();
        return bit_section + rans_section;
    }

  private:
    static void store_be32(uint8_t *p, uint32_t v) {
        p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
    }
    std::vector<uint32_t> queue_;
    std::vector<uint8_t> bits_, tail_;
    uint64_t acc_ = 0;
    uint32_t acc_bits_ = 0, ops_ = 0;
};

// Bounds-checked reader; `bad()` turns true instead of reading outside the frame.
class FrameReader {
  public:
    // returns frame size, 0 for the end-of-stream marker (ops == 0), or -1 if malformed
    int64_t begin(const uint8_t *p, size_t avail) {
        bad_ = false;
        if (avail < 4) return -1;
        ops_ = load_be32(p);
        if (ops_ == 0) return 0;
        if (avail < 12) return -1;
        uint64_t bit_section = load_be32(p + 4), rans_section = load_be32(p + 8);
        if (bit_section < 12 || rans_section < 16 || bit_section + rans_section > avail) return -1;
        bit_ptr_ = p + 12;
        bit_end_ = p + bit_section;
        rans_ptr_ = p + bit_section;
        rans_end_ = rans_ptr_ + rans_section;
        for (int k = 0; k < 4; k++) {
            x_[k] = (uint32_t)rans_ptr_[0] | ((uint32_t)rans_ptr_[1] << 8) | ((uint32_t)rans_ptr_[2] << 16) |
                    ((uint32_t)rans_ptr_[3] << 24);
            rans_ptr_ += 4;
        }
        turn_ = 0;
        acc_ = 0;
        acc_bits_ = 0;
        return (int64_t)(bit_section + rans_section);
    }
    uint32_t ops_left() const { return ops_; }
    bool bad() const { return bad_; }

    template <int BITS> int get(const Table<BITS> &t) {
        --ops_;
        uint32_t &s = x_[turn_++ & 3];
        int y = t.find(s & (kProbOne - 1));
        s = t.freq(y) * (s >> kProbBits) + (s & (kProbOne - 1)) - t.low(y);
        if (s < kRansFloor) {
            if (rans_end_ - rans_ptr_ < 2) { bad_ = true; return y; }
            s = (s << 16) + ((uint32_t)rans_ptr_[0] << 8) + rans_ptr_[1];
            rans_ptr_ += 2;
        }
        return y;
    }
    uint32_t get_raw(uint32_t nb) {
        --ops_;
        while (acc_bits_ < nb) {
            if (bit_ptr_ >= bit_end_) { bad_ = true; return 0; }
            acc_ = (acc_ << 8) | *bit_ptr_++;
            acc_bits_ += 8;
        }
        acc_bits_ -= nb;
        return (uint32_t)(acc_ >> acc_bits_) & ((1u << nb) - 1);
    }

  private:
    static uint32_t load_be32(const uint8_t *p) {
        return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    const uint8_t *bit_ptr_ = nullptr, *bit_end_ = nullptr, *rans_ptr_ = nullptr, *rans_end_ = nullptr;
    uint32_t x_[4] = {0, 0, 0, 0};
    uint32_t turn_ = 0, ops_ = 0, acc_bits_ = 0;
    uint64_t acc_ = 0;
    bool bad_ = false;
};

}  // namespace nlzm_host
#endif
