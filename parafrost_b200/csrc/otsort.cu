// parafrost_b200/csrc/otsort.cu -- segmented sort of every occurrence list by the total key
// (size, first literal, last literal, signature, clause index), hand-written; replaces
// mgpu::segmented_sort + OLIST_CMP (src/gpu/segsort.cu:37-48, src/gpu/key.cuh:67-83).
//
// The reference's comparator dereferences two clauses per comparison.  Here the 16-byte key
// of a clause is precomputed once per round (k_hist_key), gathered once per list entry, and
// the whole list is sorted on chip: in registers (<= 16 entries, one thread per list), in a
// warp's shared-memory slice (<= 512), in a CTA's shared memory (<= 8192), or - for the rare
// longer list - in place in global memory by one CTA.
// Algorithmic bytes: 4L (read entries) + 16L (key gather) + 4L (write) = 24L.
#include "common.cuh"

#define SORT_SMALL 16
#define SORT_MED 512
#define SORT_BIG 8192
#define MED_WARPS 4

struct SKey { u64 a, b; u32 id; };
__device__ __forceinline__ bool skLess(u64 a0, u64 b0, u32 i0, u64 a1, u64 b1, u32 i1) {
    if (a0 != a1) return a0 < a1;
    if (b0 != b1) return b0 < b1;
    return i0 < i1;
}
__device__ __forceinline__ void loadKey(const uint4* __restrict__ key, u32 id, u64& a, u64& b) {
    const uint4 k = key[id];
    a = ((u64)k.x << 32) | k.y;
    b = ((u64)k.z << 32) | k.w;
}

// thread per literal: sorts short lists in registers, queues the longer ones
__global__ void k_sort_small(const uint4* __restrict__ key, const u32* __restrict__ otStart, const u32* __restrict__ otSize,
                             u32* __restrict__ occurs, u32 ND, u32* __restrict__ qMed, u32* __restrict__ qBig,
                             u32* __restrict__ qHuge, DevCounters* dc) {
    for (u32 lit = 2 + blockIdx.x * blockDim.x + threadIdx.x; lit < ND; lit += gridDim.x * blockDim.x) {
        const u32 n = otSize[lit];
        if (n < 2) continue;
        if (n > SORT_SMALL) {
            if (n <= SORT_MED) qMed[atomicAdd(&dc->qMed, 1u)] = lit;
            else if (n <= SORT_BIG) qBig[atomicAdd(&dc->qBig, 1u)] = lit;
            else qHuge[atomicAdd(&dc->qHuge, 1u)] = lit;
            continue;
        }
        u32* list = occurs + otStart[lit];
        u64 ka[SORT_SMALL], kb[SORT_SMALL];
        u32 id[SORT_SMALL];
        for (u32 j = 0; j < n; j++) {
            const u32 r = list[j];
            u64 a, b;
            loadKey(key, r, a, b);
            int p = (int)j;
            while (p > 0 && skLess(a, b, r, ka[p - 1], kb[p - 1], id[p - 1])) {
                ka[p] = ka[p - 1]; kb[p] = kb[p - 1]; id[p] = id[p - 1];
                p--;
            }
            ka[p] = a; kb[p] = b; id[p] = r;
        }
        for (u32 j = 0; j < n; j++) list[j] = id[j];
    }
}

// bitonic network with mirrored first merge step: every comparator is ascending, so virtual
// +inf padding above n never has to move and arbitrary n works in place
template <typename SYNC>
__device__ __forceinline__ void bitonicShared(u64* ka, u64* kb, u32* id, u32 n, u32 tid, u32 nthreads, SYNC sync) {
    u32 P = 1;
    while (P < n) P <<= 1;
    for (u32 k = 2; k <= P; k <<= 1) {
        const u32 half = k >> 1;
        for (u32 t = tid; t < (P >> 1); t += nthreads) {
            const u32 blk = t / half, w = t - blk * half;
            const u32 i = blk * k + w, p = blk * k + k - 1 - w;
            if (p < n && skLess(ka[p], kb[p], id[p], ka[i], kb[i], id[i])) {
                u64 x = ka[i]; ka[i] = ka[p]; ka[p] = x;
                x = kb[i]; kb[i] = kb[p]; kb[p] = x;
                u32 y = id[i]; id[i] = id[p]; id[p] = y;
            }
        }
        sync();
        for (u32 j = half >> 1; j > 0; j >>= 1) {
            for (u32 t = tid; t < (P >> 1); t += nthreads) {
                const u32 i = 2 * j * (t / j) + (t % j), p = i + j;
                if (p < n && skLess(ka[p], kb[p], id[p], ka[i], kb[i], id[i])) {
                    u64 x = ka[i]; ka[i] = ka[p]; ka[p] = x;
                    x = kb[i]; kb[i] = kb[p]; kb[p] = x;
                    u32 y = id[i]; id[i] = id[p]; id[p] = y;
                }
            }
            sync();
        }
    }
}

struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// one warp per list, 17..512 entries, shared-memory slice per warp
__global__ void __launch_bounds__(MED_WARPS * 32) k_sort_med(const uint4* __restrict__ key, const u32* __restrict__ otStart,
                                                             const u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                             const u32* __restrict__ q, const DevCounters* dc) {
    __shared__ u64 ska[MED_WARPS][SORT_MED];
    __shared__ u64 skb[MED_WARPS][SORT_MED];
    __shared__ u32 sid[MED_WARPS][SORT_MED];
    const u32 w = threadIdx.x >> 5, l = threadIdx.x & 31u;
    const u32 nq = dc->qMed;
    for (u32 qi = blockIdx.x * MED_WARPS + w; qi < nq; qi += gridDim.x * MED_WARPS) {
        const u32 lit = q[qi], n = otSize[lit];
        u32* list = occurs + otStart[lit];
        for (u32 j = l; j < n; j += 32) {
            const u32 r = list[j];
            sid[w][j] = r;
            loadKey(key, r, ska[w][j], skb[w][j]);
        }
        __syncwarp();
        bitonicShared(ska[w], skb[w], sid[w], n, l, 32u, WarpSync());
        for (u32 j = l; j < n; j += 32) list[j] = sid[w][j];
        __syncwarp();
    }
}

// one CTA per list, 513..8192 entries, dynamic shared memory (160 KB)
__global__ void __launch_bounds__(512) k_sort_big(const uint4* __restrict__ key, const u32* __restrict__ otStart,
                                                  const u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                  const u32* __restrict__ q, const DevCounters* dc) {
    extern __shared__ u64 dyn[];
    u64* ska = dyn;
    u64* skb = dyn + SORT_BIG;
    u32* sid = (u32*)(dyn + 2 * SORT_BIG);
    const u32 nq = dc->qBig;
    for (u32 qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        const u32 lit = q[qi], n = otSize[lit];
        u32* list = occurs + otStart[lit];
        for (u32 j = threadIdx.x; j < n; j += blockDim.x) {
            const u32 r = list[j];
            sid[j] = r;
            loadKey(key, r, ska[j], skb[j]);
        }
        __syncthreads();
        bitonicShared(ska, skb, sid, n, threadIdx.x, blockDim.x, BlockSync());
        for (u32 j = threadIdx.x; j < n; j += blockDim.x) list[j] = sid[j];
        __syncthreads();
    }
}

// one CTA per list, > 8192 entries: same network in place in global memory, keys gathered per compare
__global__ void __launch_bounds__(1024) k_sort_huge(const uint4* __restrict__ key, const u32* __restrict__ otStart,
                                                    const u32* __restrict__ otSize, u32* occurs,
                                                    const u32* __restrict__ q, const DevCounters* dc) {
    const u32 nq = dc->qHuge;
    for (u32 qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        const u32 lit = q[qi], n = otSize[lit];
        volatile u32* list = occurs + otStart[lit];
        u32 P = 1;
        while (P < n) P <<= 1;
        for (u32 k = 2; k <= P; k <<= 1) {
            const u32 half = k >> 1;
            for (u32 j = half, first = 1; j > 0; j >>= 1, first = 0) {
                for (u32 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                    u32 i, p;
                    if (first) { const u32 blk = t / half, w = t - blk * half; i = blk * k + w; p = blk * k + k - 1 - w; }
                    else { i = 2 * j * (t / j) + (t % j); p = i + j; }
                    if (p < n) {
                        const u32 ri = list[i], rp = list[p];
                        u64 ai, bi, ap, bp;
                        loadKey(key, ri, ai, bi);
                        loadKey(key, rp, ap, bp);
                        if (skLess(ap, bp, rp, ai, bi, ri)) { list[i] = rp; list[p] = ri; }
                    }
                }
                __syncthreads();
            }
        }
    }
}

__global__ void k_sort_reset(DevCounters* dc) { dc->qMed = 0; dc->qBig = 0; dc->qHuge = 0; }

void launchSortOT(Ctx* c) {
    const size_t bigSmem = (size_t)SORT_BIG * (8 + 8 + 4);
    if (!c->attrSort) {
        cudaFuncSetAttribute(k_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bigSmem);
        c->attrSort = true;
    }
    LAUNCH(c, k_sort_reset, 1, 1, 0, c->dc);
    LAUNCH(c, k_sort_small, gridFor(c->ND, 128), 128, 0, c->key, c->otStart, c->otSize, c->occurs, c->ND, c->qMed, c->qBig, c->qHuge, c->dc);
    LAUNCH(c, k_sort_med, 148 * 4, MED_WARPS * 32, 0, c->key, c->otStart, c->otSize, c->occurs, c->qMed, c->dc);
    LAUNCH(c, k_sort_big, 148, 512, bigSmem, c->key, c->otStart, c->otSize, c->occurs, c->qBig, c->dc);
    LAUNCH(c, k_sort_huge, 32, 1024, 0, c->key, c->otStart, c->otSize, c->occurs, c->qHuge, c->dc);
}
