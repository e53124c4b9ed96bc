// Stream reader under AddressSanitizer + UBSan (compiled and run by tests/test_host_units.py):
// a good stream must decode to the expected bytes; truncations and bit flips must be rejected or
// decode to something else without touching memory outside the buffers.
//   decoder_sanitize <stream.nlzm> <expected.bin>
#include "../include/nlzm_codec.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static std::vector<uint8_t> slurp(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { printf("cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v(n);
    if (n && fread(v.data(), 1, n, f) != n) exit(2);
    fclose(f);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::vector<uint8_t> s = slurp(argv[1]), want = slurp(argv[2]);
    uint8_t *out = nullptr;
    uint64_t n = 0;
    int rc = nlzm_codec_decompress(s.data(), s.size(), &out, &n);
    if (rc || n != want.size() || memcmp(out, want.data(), n)) { printf("good stream: rc=%d n=%llu\n", rc, (unsigned long long)n); return 1; }
    nlzm_codec_free(out);
    unsigned rejected = 0, decoded = 0;
    for (size_t cut = 0; cut < s.size(); cut += 211) {
        std::vector<uint8_t> b(s.begin(), s.begin() + cut);          // exact-size copy: reads past `cut` are caught
        out = nullptr;
        rc = nlzm_codec_decompress(b.data(), b.size(), &out, &n);
        rc ? ++rejected : ++decoded;
        nlzm_codec_free(out);
    }
    srand(7);
    for (int i = 0; i < 400; i++) {
        std::vector<uint8_t> b = s;
        b[rand() % b.size()] ^= (uint8_t)(1 << (rand() % 8));
        out = nullptr;
        rc = nlzm_codec_decompress(b.data(), b.size(), &out, &n);
        rc ? ++rejected : ++decoded;
        nlzm_codec_free(out);
    }
    printf("decoder sanitize ok: %u rejected, %u decoded\n", rejected, decoded);
    return 0;
}
