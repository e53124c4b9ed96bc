// TEST INFRASTRUCTURE: a few CUDA runtime names mapped onto the host heap so that engine.cu can be
// compiled by g++ with -DNLZM_EMU and its kernel bodies exercised sequentially without a GPU.
// Never part of the product library.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <chrono>

enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuEvent *cudaEvent_t;

static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "ok"; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent(); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return 0;
}
static inline cudaError_t cudaGetLastError() { return 0; }
#define cudaEventDisableTiming 0
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, int) { return cudaEventCreate(e); }
