// Device-wide primitives the stages are built from: LSD radix sort of (key, value) pairs,
// exclusive sum, inclusive max, sum. CUDA build: CUB (header library shipped with the toolkit),
// all on the caller's stream with caller-provided temporary storage. Emulation build (tests
// only): the obvious sequential versions.
#pragma once
#include "platform.cuh"

#ifndef NLZM_EMU
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>
#include <cuda/functional>

struct PrimTemp {
    void *ptr = nullptr;
    size_t bytes = 0;
};

#define NLZM_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t e_ = (expr);                              \
        if (e_ != cudaSuccess) return (int)e_;                \
    } while (0)

// size query helpers (worst case over the calls the engine makes for n items)
static inline size_t prim_temp_bytes(u64 n) {
    size_t need = 0, t = 0;
    cub::DoubleBuffer<u64> k64(nullptr, nullptr);
    cub::DoubleBuffer<u32> k32(nullptr, nullptr), v32(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, t, k64, v32, (i64)n, 0, 64);
    need = t > need ? t : need;
    cub::DeviceRadixSort::SortPairs(nullptr, t, k32, v32, (i64)n, 0, 32);
    need = t > need ? t : need;
    cub::DeviceScan::ExclusiveSum(nullptr, t, (u32 *)nullptr, (u32 *)nullptr, (i64)n);
    need = t > need ? t : need;
    cub::DeviceScan::InclusiveScan(nullptr, t, (u32 *)nullptr, (u32 *)nullptr, ::cuda::maximum<>{}, (i64)n);
    need = t > need ? t : need;
    cub::DeviceReduce::Sum(nullptr, t, (u32 *)nullptr, (u64 *)nullptr, (i64)n);
    need = t > need ? t : need;
    return need + 256;
}

// Sorts pairs by key bits [begin_bit, end_bit). Result lands in buffer `*sel` (0: k0/v0, 1: k1/v1).
static inline int prim_sort_pairs64(PrimTemp &tmp, u64 *k0, u64 *k1, u32 *v0, u32 *v1, u64 n, int begin_bit,
                                    int end_bit, cudaStream_t st, int *sel) {
    cub::DoubleBuffer<u64> k(k0, k1);
    cub::DoubleBuffer<u32> v(v0, v1);
    size_t t = tmp.bytes;
    nlzm_launch_begin("cub_radix_sort_pairs", st);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp.ptr, t, k, v, (i64)n, begin_bit, end_bit, st);
    nlzm_launch_end(st);
    if (e != cudaSuccess) return (int)e;
    *sel = k.selector;
    return 0;
}
static inline int prim_sort_pairs32(PrimTemp &tmp, u32 *k0, u32 *k1, u32 *v0, u32 *v1, u64 n, int begin_bit,
                                    int end_bit, cudaStream_t st, int *sel) {
    cub::DoubleBuffer<u32> k(k0, k1);
    cub::DoubleBuffer<u32> v(v0, v1);
    size_t t = tmp.bytes;
    nlzm_launch_begin("cub_radix_sort_pairs", st);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp.ptr, t, k, v, (i64)n, begin_bit, end_bit, st);
    nlzm_launch_end(st);
    if (e != cudaSuccess) return (int)e;
    *sel = k.selector;
    return 0;
}
static inline int prim_exclusive_sum(PrimTemp &tmp, const u32 *in, u32 *out, u64 n, cudaStream_t st) {
    size_t t = tmp.bytes;
    nlzm_launch_begin("cub_exclusive_sum", st);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp.ptr, t, in, out, (i64)n, st);
    nlzm_launch_end(st);
    if (e != cudaSuccess) return (int)e;
    return 0;
}
static inline int prim_inclusive_max(PrimTemp &tmp, const u32 *in, u32 *out, u64 n, cudaStream_t st) {
    size_t t = tmp.bytes;
    nlzm_launch_begin("cub_inclusive_max", st);
    cudaError_t e = cub::DeviceScan::InclusiveScan(tmp.ptr, t, in, out, ::cuda::maximum<>{}, (i64)n, st);
    nlzm_launch_end(st);
    if (e != cudaSuccess) return (int)e;
    return 0;
}
static inline int prim_sum(PrimTemp &tmp, const u32 *in, u64 *out, u64 n, cudaStream_t st) {
    size_t t = tmp.bytes;
    nlzm_launch_begin("cub_reduce_sum", st);
    cudaError_t e = cub::DeviceReduce::Sum(tmp.ptr, t, in, out, (i64)n, st);
    nlzm_launch_end(st);
    if (e != cudaSuccess) return (int)e;
    return 0;
}
#else
#include <algorithm>
#include <vector>
struct PrimTemp {
    void *ptr = nullptr;
    size_t bytes = 0;
};
static inline size_t prim_temp_bytes(u64) { return 256; }
template <class K>
static inline int prim_sort_emu(K *k0, K *k1, u32 *v0, u32 *v1, u64 n, int bb, int eb, int *sel) {
    std::vector<u64> idx(n);
    for (u64 i = 0; i < n; i++) idx[i] = i;
    const int nb = eb - bb;
    const u64 mask = nb >= 64 ? ~0ull : ((1ull << nb) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](u64 a, u64 b) {
        return (((u64)k0[a] >> bb) & mask) < (((u64)k0[b] >> bb) & mask);
    });
    for (u64 i = 0; i < n; i++) { k1[i] = k0[idx[i]]; v1[i] = v0[idx[i]]; }
    *sel = 1;
    return 0;
}
static inline int prim_sort_pairs64(PrimTemp &, u64 *k0, u64 *k1, u32 *v0, u32 *v1, u64 n, int bb, int eb,
                                    cudaStream_t, int *sel) { return prim_sort_emu(k0, k1, v0, v1, n, bb, eb, sel); }
static inline int prim_sort_pairs32(PrimTemp &, u32 *k0, u32 *k1, u32 *v0, u32 *v1, u64 n, int bb, int eb,
                                    cudaStream_t, int *sel) { return prim_sort_emu(k0, k1, v0, v1, n, bb, eb, sel); }
static inline int prim_exclusive_sum(PrimTemp &, const u32 *in, u32 *out, u64 n, cudaStream_t) {
    u32 s = 0;
    for (u64 i = 0; i < n; i++) { u32 v = in[i]; out[i] = s; s += v; }
    return 0;
}
static inline int prim_inclusive_max(PrimTemp &, const u32 *in, u32 *out, u64 n, cudaStream_t) {
    u32 s = 0;
    for (u64 i = 0; i < n; i++) { s = in[i] > s ? in[i] : s; out[i] = s; }
    return 0;
}
static inline int prim_sum(PrimTemp &, const u32 *in, u64 *out, u64 n, cudaStream_t) {
    u64 s = 0;
    for (u64 i = 0; i < n; i++) s += in[i];
    *out = s;
    return 0;
}
#endif
