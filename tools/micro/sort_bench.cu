// micro-benchmark: which CUB radix sort variant is fastest for the candidate-tuple sort
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
typedef unsigned long long u64; typedef unsigned u32;
__global__ void gen(u64 *k, u32 *v32, u64 n, int layout) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    u64 h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    u32 pos = (u32)(h % 100000000u); u32 len = 4 + (u32)((h >> 40) % 9); u32 dist = (u32)((h >> 20) & 0xFFFFFF);
    if (layout == 0) { k[i] = ((u64)pos << 37) | ((u64)len << 28) | dist; }
    else if (layout == 1) { k[i] = ((u64)pos << 9) | len; v32[i] = dist; }
    else { k[i] = ((u64)dist << 36) | ((u64)pos << 9) | len; }
}
static u64 n = 437000000ull; static u64 *k0, *k1; static u32 *v0, *v1; static void *tmp; static size_t tb = 1ull << 28;
static void G(int l) { gen<<<(unsigned)((n + 255) / 256), 256>>>(k0, v0, n, l); cudaDeviceSynchronize(); }
static void keys(int bb, int eb) { cub::DoubleBuffer<u64> kb(k0, k1); size_t t = tb; cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, t, kb, (long)n, bb, eb); if (e) printf("err %d\n", (int)e); }
static void pairs(int bb, int eb) { cub::DoubleBuffer<u64> kb(k0, k1); cub::DoubleBuffer<u32> vb(v0, v1); size_t t = tb; cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, t, kb, vb, (long)n, bb, eb); if (e) printf("err %d\n", (int)e); }
static void pairs32(int bb, int eb) { cub::DoubleBuffer<u32> kb((u32 *)k0, (u32 *)k1); cub::DoubleBuffer<u32> vb(v0, v1); size_t t = tb; cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, t, kb, vb, (long)n, bb, eb); if (e) printf("err %d\n", (int)e); }
__global__ void gen32(u32 *k, u32 *v, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    u64 h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    u32 pos = (u32)(h % 100000000u); u32 len = 4 + (u32)((h >> 40) % 9); u32 dist = (u32)((h >> 20) & 0xFFFFFF);
    k[i] = (pos << 5) | (len >> 4); v[i] = ((len & 15) << 28) | dist;
}
template <class F> static float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaEventRecord(a); f(); f(); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3; }
int main() {
    cudaMalloc(&k0, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4); cudaMalloc(&tmp, tb);
    G(0); printf("keys64 bits[28,64)         %.2f ms\n", timeit([] { keys(28, 64); }));
    G(1); printf("pairs64+32 bits[0,36)      %.2f ms\n", timeit([] { pairs(0, 36); }));
    G(2); printf("keys64 bits[0,36) lowkey   %.2f ms\n", timeit([] { keys(0, 36); }));
    G(2); printf("keys64 bits[9,36) pos only %.2f ms\n", timeit([] { keys(9, 36); }));
    G(0); printf("keys64 bits[32,64) 4 pass  %.2f ms\n", timeit([] { keys(32, 64); }));
    G(0); printf("keys64 bits[37,64) pos     %.2f ms\n", timeit([] { keys(37, 64); }));
    gen32<<<(unsigned)((n + 255) / 256), 256>>>((u32 *)k0, v0, n); cudaDeviceSynchronize();
    printf("pairs32+32 bits[0,32)      %.2f ms\n", timeit([] { pairs32(0, 32); }));
    gen32<<<(unsigned)((n + 255) / 256), 256>>>((u32 *)k0, v0, n); cudaDeviceSynchronize();
    printf("pairs32+32 bits[5,32)      %.2f ms\n", timeit([] { pairs32(5, 32); }));
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
