// host_bench — times the host pipeline (parser + stream writer) alone, over recorded candidates.
// TEST / MEASUREMENT TOOL: the candidate file is produced by tools/host_bench.py from the CPU oracle,
// so this never ships and never stands in for the engine.
//   host_bench <text> <offsets.u64> <dist.u32> <len.u16> <window_bits> [out.nlzm] [repeats]
#include "../include/nlzm_mf.h"
#include "../nlzm_b200/csrc/host/pipeline.hpp"
#include <chrono>
#include <stdio.h>
#include <stdlib.h>

using namespace nlzm_host;

template <class T> static std::vector<T> slurp(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T) + 1);
    if (n && fread(v.data(), 1, n, f) != n) { fprintf(stderr, "short read %s\n", path); exit(2); }
    fclose(f);
    v.resize(n / sizeof(T));
    return v;
}

struct ReplayFinders {
    const uint64_t *off;
    const uint32_t *dist;
    const uint16_t *len;
    uint64_t served = 0;
    template <class T> void FindAndUpdate(T &st, uint64_t a) {
        const uint32_t *d = dist + off[a];
        const uint16_t *l = len + off[a];
        st.merge_steps((uint32_t)(off[a + 1] - off[a]), [d](uint32_t j) { return nlzm_host::candidate_entry(d[j], nlzm_host::candidate_code(d[j])); }, [l](uint32_t j) { return (uint32_t)l[j]; });
        served += off[a + 1] - off[a];
    }
};

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: host_bench text offsets dist len window_bits [out] [repeats]\n"); return 2; }
    auto text = slurp<uint8_t>(argv[1]);
    auto off = slurp<uint64_t>(argv[2]);
    auto dist = slurp<uint32_t>(argv[3]);
    auto len = slurp<uint16_t>(argv[4]);
    const uint64_t n = text.size();
    text.resize(n + 16);
    nlzm_mf_geometry g;
    nlzm_mf_get_geometry(n, (uint32_t)atoi(argv[5]), &g);
    int repeats = argc > 7 ? atoi(argv[7]) : 1;
    double best = 1e30;
    std::vector<uint8_t> out;
    EncodeCounters ec;
    for (int r = 0; r < repeats; r++) {
        ReplayFinders rf{off.data(), dist.data(), len.data()};
        out.clear();
        out.push_back((uint8_t)(g.hist_bits >> 8)); out.push_back((uint8_t)g.hist_bits);
        out.push_back((uint8_t)(g.frame_bits >> 8)); out.push_back((uint8_t)g.frame_bits);
        ec = EncodeCounters();
        auto t0 = std::chrono::steady_clock::now();
        encode_stream(text.data(), n, g.hist_bits, g.chunk_size, g.feed_size, rf, out, ec);
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best) best = s;
        out.insert(out.end(), 4, 0);
    }
    printf("{\"bytes\": %llu, \"stream\": %zu, \"seconds\": %.4f, \"MBps\": %.3f, \"parses\": %llu, \"literals\": %llu, "
           "\"matches\": %llu, \"reps\": %llu}\n", (unsigned long long)n, out.size(), best, n / best / 1e6,
           (unsigned long long)ec.parses, (unsigned long long)ec.literals, (unsigned long long)ec.matches,
           (unsigned long long)ec.reps);
    if (argc > 6 && argv[6][0]) {
        FILE *f = fopen(argv[6], "wb");
        fwrite(out.data(), 1, out.size(), f);
        fclose(f);
    }
    return 0;
}
