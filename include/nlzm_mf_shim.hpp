// nlzm_mf_shim.hpp — C++ host shim: the reference's finder interface on top of the C ABI.
//
// The reference encoder (nauful/NLZM 1.03) drives four finder objects by name:
//     mf_mem += X.Init(...)                                   NLZM.cpp:1749-1753
//     X.FindAndUpdate(mt, hash, dict.hist_pos + p, dict)      NLZM.cpp:1516-1540 (per position)
//     X.Shift(window_size)                                    NLZM.cpp:1786-1792
//     X.Release()                                             NLZM.cpp:1901-1904
// GpuMatchFinders keeps those four verbs. One object replaces ht2 + ht3 + bt4 + rk: its
// FindAndUpdate(mt, abs_pos) issues exactly the mt.Update(dist, len) calls the four reference
// finders would have issued together at that position (R2 semantics, see nlzm_mf.h), reading them
// from candidate blocks that the GPU produces ahead of the parser: while the host parses and codes
// block N (slot N&1), block N+1 is already being computed (submit/fetch double buffering).
//
// Error behaviour mirrors the reference's ASSERT (NLZM.cpp:25): print and exit(-1).
#ifndef NLZM_MF_SHIM_HPP
#define NLZM_MF_SHIM_HPP

#include "nlzm_mf.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

struct GpuMatchFinders {
    nlzm_mf *mf = nullptr;
    uint64_t flen = 0, block = 0;
    uint64_t cur_begin = 0, cur_end = 0;     // block currently readable on the host
    uint64_t next_begin = 0;                 // first position not yet submitted
    int cur_slot = 0;
    bool next_pending = false;
    nlzm_mf_view view{};
    uint64_t steps_served = 0, blocks_fetched = 0;

    static void die(nlzm_mf *h, const char *what, int rc) {
        printf("Assert failed nlzm_mf %s rc=%d: %s\n", what, rc, nlzm_mf_last_error(h));
        exit(-1);
    }

    // data: the whole input (it is kept resident in HBM); block_len: positions per GPU call.
    // Returns the window size in bytes (the reference's Init returns bytes allocated for a printf).
    uint32_t Init(uint32_t hist_bits, const uint8_t *data, uint64_t file_len, int device = 0, uint64_t block_len = 0,
                  uint32_t finder_mask = NLZM_MF_ALL) {
        nlzm_mf_geometry g;
        nlzm_mf_get_geometry(file_len, hist_bits, &g);
        if (block_len == 0) block_len = g.window > (32u << 20) ? g.window : (32u << 20);
        if (block_len > (1ull << 28)) block_len = 1ull << 28;
        nlzm_mf_config cfg{};
        cfg.struct_size = sizeof cfg;
        cfg.hist_bits = hist_bits;
        cfg.file_len = file_len;
        cfg.device = device;
        cfg.finder_mask = finder_mask;
        cfg.max_range = block_len;
        int rc = nlzm_mf_create(&cfg, &mf);
        if (rc) die(nullptr, "create", rc);
        rc = nlzm_mf_set_input(mf, data, file_len);
        if (rc) die(mf, "set_input", rc);
        flen = file_len;
        block = block_len;
        cur_begin = cur_end = next_begin = 0;
        cur_slot = 1;
        next_pending = false;
        steps_served = blocks_fetched = 0;
        submit_next();
        return g.window;
    }

    void submit_next() {
        if (next_begin >= flen) return;
        uint64_t e = next_begin + block < flen ? next_begin + block : flen;
        int rc = nlzm_mf_submit(mf, next_begin, e, cur_slot ^ 1);
        if (rc) die(mf, "submit", rc);
        next_begin = e;
        next_pending = true;
    }

    void advance() {
        if (!next_pending) { printf("Assert failed nlzm_mf position past the end of the input\n"); exit(-1); }
        cur_slot ^= 1;
        int rc = nlzm_mf_fetch(mf, cur_slot, &view);
        if (rc) die(mf, "fetch", rc);
        next_pending = false;
        cur_begin = view.begin;
        cur_end = view.end;
        ++blocks_fetched;
        submit_next();                       // GPU works on block N+1 while the host consumes block N
    }

    // MatchTableT needs Update(uint32 dist, uint16 len) — NLZM.cpp:835-852. Positions must not go
    // backwards by more than the current block (the parser visits positions in ascending order).
    template <class MatchTableT> void FindAndUpdate(MatchTableT &mt, uint64_t abs_pos) {
        while (abs_pos >= cur_end) advance();
        if (abs_pos < cur_begin) { printf("Assert failed nlzm_mf position %llu before the current block\n", (unsigned long long)abs_pos); exit(-1); }
        const uint64_t i = abs_pos - cur_begin;
        const uint32_t b = view.offsets[i], e = view.offsets[i + 1];
        for (uint32_t s = b; s < e; s++) mt.Update(NLZM_MF_STEP_DIST(view.steps[s]), (uint16_t)NLZM_MF_STEP_LEN(view.steps[s]));
        steps_served += e - b;
    }

    void Shift(uint32_t) {}                  // ring shifts are part of the engine's closed-form geometry

    void Release() {
        if (mf) {
            if (next_pending) { nlzm_mf_view v; nlzm_mf_fetch(mf, cur_slot ^ 1, &v); next_pending = false; }
            nlzm_mf_destroy(mf);
        }
        mf = nullptr;
    }
};

#endif
