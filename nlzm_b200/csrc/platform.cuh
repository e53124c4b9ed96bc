// Build-mode glue. The product is compiled by nvcc for sm_100a. The same kernel bodies can also
// be compiled by g++ with -DNLZM_EMU into a *test-only* library (tests/emu) in which every
// "kernel" is a sequential loop over its thread index: that library exists so that the kernel
// logic can be unit-tested in a container without a GPU. It is never shipped or loaded by the
// nlzm_b200 package (nlzm_b200/_lib.py loads libnlzm_mf.so only and fails loudly without CUDA).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#ifndef NLZM_EMU
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__

// One thread per element i < n; the kernel is named k_<name> so that ncu shows a readable symbol.
#define NLZM_KERNEL_1D(name, ParamsT)                                                   \
    __global__ void __launch_bounds__(256) k_##name(const ParamsT p, u64 n) {           \
        u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;                             \
        if (i < n) name##_body(p, i);                                                   \
    }                                                                                   \
    static inline void launch_##name(const ParamsT &p, u64 n, cudaStream_t st) {        \
        if (n == 0) return;                                                             \
        nlzm_launch_begin("k_" #name, st);                                              \
        k_##name<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n);                    \
        nlzm_launch_end(st);                                                            \
    }
// One CTA of NT threads runs name##_cta(p, block, thread, smem); dynamic shared memory.
#define NLZM_KERNEL_CTA(name, ParamsT, NT) NLZM_KERNEL_CTA_OCC(name, ParamsT, NT, 1)
#define NLZM_KERNEL_CTA_OCC(name, ParamsT, NT, MINB)                                    \
    __global__ void __launch_bounds__(NT, MINB) k_##name(const ParamsT p) {             \
        extern __shared__ __align__(16) u8 nlzm_smem[];                                 \
        name##_cta(p, blockIdx.x, threadIdx.x, nlzm_smem);                              \
    }                                                                                   \
    static inline int launch_##name(const ParamsT &p, u64 grid, size_t smem, cudaStream_t st) { \
        if (grid == 0) return 0;                                                        \
        if (smem > 48 * 1024) {                                                         \
            int e = nlzm_smem_opt_in((const void *)k_##name, smem);                     \
            if (e) return e;                                                            \
        }                                                                               \
        nlzm_launch_begin("k_" #name, st);                                              \
        k_##name<<<(unsigned)grid, NT, smem, st>>>(p);                                  \
        nlzm_launch_end(st);                                                            \
        return (int)cudaGetLastError();                                                 \
    }
#define NLZM_CTA_SYNC() __syncthreads()
// Opt-in to more than 48 KiB of dynamic shared memory. The attribute belongs to the (kernel, device) pair, and one
// process may drive engines on several devices, so the size already granted is remembered per device.
#include <atomic>
static inline int nlzm_smem_opt_in(const void *kernel, size_t smem) {
    struct Granted { std::atomic<const void *> fn{nullptr}; std::atomic<size_t> bytes[64]; };
    static Granted table[32];                                   // a handful of kernels need the opt-in
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
    Granted *g = nullptr;
    for (auto &t : table) {
        const void *cur = t.fn.load(std::memory_order_acquire);
        if (cur == kernel) { g = &t; break; }
        if (cur == nullptr) {
            const void *expect = nullptr;
            if (t.fn.compare_exchange_strong(expect, kernel, std::memory_order_acq_rel) || expect == kernel) { g = &t; break; }
        }
    }
    if (g && dev >= 0 && g->bytes[dev].load(std::memory_order_acquire) >= smem) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (g && dev >= 0) {                                        // racing threads set the same attribute: benign
        size_t old = g->bytes[dev].load(std::memory_order_relaxed);
        while (old < smem && !g->bytes[dev].compare_exchange_weak(old, smem, std::memory_order_release)) {}
    }
    return 0;
}
// ---- bulk asynchronous copy global -> shared through the TMA unit (cp.async.bulk, SASS UBLKCP), completion counted
//      in bytes on an mbarrier. One elected thread issues the copies of a tile; everybody waits on the barrier.
DEV u32 nlzm_smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
DEV void nlzm_mbar_init(u32 bar, u32 arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void nlzm_mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
DEV void nlzm_bulk_g2s(u32 dst_smem, const void *src_gmem, u32 bytes, u32 bar) {      // 16-byte aligned, bytes % 16 == 0
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
DEV void nlzm_mbar_wait(u32 bar, u32 parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar), "r"(parity) : "memory");
}
DEV u32 nlzm_atomic_add(u32 *p, u32 v) { return atomicAdd(p, v); }
DEV u64 nlzm_atomic_add64(unsigned long long *p, u64 v) { return atomicAdd(p, (unsigned long long)v); }
DEV u32 nlzm_atomic_max(u32 *p, u32 v) { return atomicMax(p, v); }
DEV u32 nlzm_atomic_min(u32 *p, u32 v) { return atomicMin(p, v); }
HD int nlzm_ctz64(u64 v) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
#else
// ---- emulation (tests only) ----
#include <stdlib.h>
#include <stdio.h>
#define HD inline
#define DEV inline
#define __global__
#define __device__
#define __host__
#define __restrict__
#define __forceinline__
typedef int cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
#define NLZM_KERNEL_1D(name, ParamsT)                                                   \
    static inline void launch_##name(const ParamsT &p, u64 n, cudaStream_t st) {        \
        nlzm_launch_begin("k_" #name, st);                                              \
        for (u64 i = 0; i < n; i++) name##_body(p, i);                                  \
        nlzm_launch_end(st);                                                            \
    }
// CTA kernels run with real host threads and a barrier, so the atomics are real too
#include <barrier>
#include <thread>
#include <vector>
struct EmuCta { std::barrier<> *bar; };
extern thread_local EmuCta nlzm_emu_cta;
#define NLZM_CTA_SYNC() nlzm_emu_cta.bar->arrive_and_wait()
#define NLZM_KERNEL_CTA_OCC(name, ParamsT, NT, MINB) NLZM_KERNEL_CTA(name, ParamsT, NT)
#define NLZM_KERNEL_CTA(name, ParamsT, NT)                                              \
    static inline int launch_##name(const ParamsT &p, u64 grid, size_t smem, cudaStream_t st) { \
        if (grid == 0) return 0;                                                        \
        nlzm_launch_begin("k_" #name, st);                                              \
        std::vector<u8> sm(smem + 16);                                                  \
        std::barrier<> bar(NT);                                                         \
        std::vector<std::thread> th;                                                    \
        for (u32 t = 0; t < NT; t++)                                                    \
            th.emplace_back([&, t]() {                                                  \
                nlzm_emu_cta.bar = &bar;                                                \
                for (u64 b = 0; b < grid; b++) {                                        \
                    name##_cta(p, (u32)b, t, sm.data());                                \
                    bar.arrive_and_wait();                                              \
                }                                                                       \
            });                                                                         \
        for (auto &x : th) x.join();                                                    \
        nlzm_launch_end(st);                                                            \
        return 0;                                                                       \
    }
static inline u32 nlzm_atomic_add(u32 *p, u32 v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline u64 nlzm_atomic_add64(unsigned long long *p, u64 v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline u32 nlzm_atomic_max(u32 *p, u32 v) {
    u32 o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline u32 nlzm_atomic_min(u32 *p, u32 v) {
    u32 o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline int nlzm_ctz64(u64 v) { return __builtin_ctzll(v); }
#endif

// launch accounting (engine.cu): every kernel launch is counted; with profiling on, it is also
// bracketed by CUDA events on its stream and its time accumulated per kernel name.
void nlzm_launch_begin(const char *name, cudaStream_t st);
void nlzm_launch_end(cudaStream_t st);
