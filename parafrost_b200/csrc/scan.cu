// parafrost_b200/csrc/scan.cu -- device-wide exclusive scans (hand-written; replaces the
// cub::DeviceScan calls at src/gpu/memory.cu:429-430, elimination.cu:85-92, recycle.cu:84-93).
//
// Three launches: per-tile reduce -> single-CTA scan of the tile sums -> per-tile downsweep
// (one launch for short ranges, k_scan_small).
// Algorithmic bytes: read n + write n (+ the tile sums); HBM-bound.
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <typename T>
__device__ __forceinline__ T blockExclusive(T v, T* smem /* >= 32 */, T& total) {
    // inclusive warp scan
    const u32 l = threadIdx.x & 31u, w = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, x, o); if (l >= (u32)o) x += t; }
    if (l == 31) smem[w] = x;
    __syncthreads();
    if (w == 0) {
        const u32 nw = blockDim.x >> 5;
        T s = (l < nw) ? smem[l] : T(0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, s, o); if (l >= (u32)o) s += t; }
        smem[l] = s;  // inclusive over warps
    }
    __syncthreads();
    const T warpBase = w ? smem[w - 1] : T(0);
    total = smem[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warpBase + x - v;
}

template <typename T>
__global__ void k_scan_reduce(const T* __restrict__ in, u64 n, T* __restrict__ tileSums) {
    __shared__ T sm[32];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += in[base + k];
    T total;
    blockExclusive<T>(s, sm, total);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

template <typename T>
__global__ void k_scan_tiles(T* tileSums, u32 ntiles, T init, T* totalOut) {
    __shared__ T sm[32];
    __shared__ T carry;
    if (threadIdx.x == 0) carry = init;
    __syncthreads();
    for (u32 b = 0; b < ntiles; b += blockDim.x) {
        const u32 i = b + threadIdx.x;
        const T v = (i < ntiles) ? tileSums[i] : T(0);
        T total;
        const T ex = blockExclusive<T>(v, sm, total);
        const T c0 = carry;
        if (i < ntiles) tileSums[i] = c0 + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c0 + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && totalOut) *totalOut = carry;
}

template <typename T>
__global__ void k_scan_down(const T* in, T* out, u64 n, const T* __restrict__ tileSums) {
    __shared__ T sm[32];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : T(0); s += v[k]; }
    T total;
    T ex = blockExclusive<T>(s, sm, total) + tileSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

// Short ranges (worklists, radix digit tables, elected variables: most scans of a round) in ONE launch:
// a single CTA walks the range in chunks of 1024 x 4 elements with a running carry.
#define SCAN_SMALL_MAX (32u << 10)
template <typename T>
__global__ void __launch_bounds__(1024) k_scan_small(const T* in, T* out, u32 n, T init, T* totalOut) {
    __shared__ T sm[32];
    __shared__ T carry;
    if (threadIdx.x == 0) carry = init;
    __syncthreads();
    for (u32 b = 0; b < n; b += 4096) {
        const u32 i0 = b + threadIdx.x * 4;
        T v[4]; T s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (i0 + k < n) ? in[i0 + k] : T(0); s += v[k]; }
        T total;
        T ex = blockExclusive<T>(s, sm, total) + carry;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i0 + k < n) out[i0 + k] = ex; ex += v[k]; }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && totalOut) *totalOut = carry;
}

template <typename T>
static void scanImpl(Ctx* c, const T* in, T* out, u64 n, T init, T* tmp, T* totalOut) {
    if (!n) return;  // callers never scan empty ranges with a total
    if (n <= SCAN_SMALL_MAX) { LAUNCH(c, k_scan_small<T>, 1, 1024, 0, in, out, (u32)n, init, totalOut); KB(c, 2.0 * sizeof(T) * n); return; }
    const u32 ntiles = divup(n, SCAN_TILE);
    LAUNCH(c, k_scan_reduce<T>, ntiles, SCAN_THREADS, 0, in, n, tmp);
    KB(c, (double)sizeof(T) * n);
    LAUNCH(c, k_scan_tiles<T>, 1, 1024, 0, tmp, ntiles, init, totalOut);
    LAUNCH(c, k_scan_down<T>, ntiles, SCAN_THREADS, 0, in, out, n, tmp);
    KB(c, 2.0 * sizeof(T) * n);
}

void scanExclusiveU32(Ctx* c, const u32* in, u32* out, u64 n, u32 init, u32* totalOut) {
    scanImpl<u32>(c, in, out, n, init, c->scanTmp, totalOut);
}
void scanExclusiveU64(Ctx* c, const u64* in, u64* out, u64 n, u64 init) {
    scanImpl<u64>(c, in, out, n, init, c->scanTmp64, (u64*)nullptr);
}
