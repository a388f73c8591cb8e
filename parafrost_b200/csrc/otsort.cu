// parafrost_b200/csrc/otsort.cu -- segmented sort of every occurrence list by the total key
// (size, first literal, last literal, signature, clause index), hand-written; replaces
// mgpu::segmented_sort + OLIST_CMP (src/gpu/segsort.cu:37-48, src/gpu/key.cuh:67-83).
//
// The reference's comparator dereferences two clauses per comparison.  Here the 16-byte key
// of a clause is precomputed once per round (k_ot_count), gathered once per list entry, and
// the whole list is sorted on chip: in registers (<= 16 entries, one thread per list), in a
// warp's shared-memory slice (<= 512), in a CTA's shared memory (<= 8192), or - for the rare
// longer list - in place in global memory by one CTA.
// Algorithmic bytes: 4L (read entries) + 16L (key gather) + 4L (write) = 24L.
#include "common.cuh"

#define SORT_MED 512
#define SORT_BIG 8192
#define MED_WARPS 4
#define NCLASS 9   // list-length classes: <=4 <=8 <=16 <=32 <=64 <=128 (registers), <=512 (warp smem), <=8192 (CTA smem), longer

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

// ------------------------------------------------------------------ length classes
__device__ __forceinline__ int sortClass(u32 n) {
    if (n < 2) return -1;
    if (n <= 4) return 0;
    if (n <= 8) return 1;
    if (n <= 16) return 2;
    if (n <= 32) return 3;
    if (n <= 64) return 4;
    if (n <= 128) return 5;
    if (n <= SORT_MED) return 6;
    if (n <= SORT_BIG) return 7;
    return 8;
}
// pass 1: class sizes (block-reduced, one atomic per class and block)
// The sort can be restricted to the lists somebody will read in order:
//   vinfo (lcve.cu election words): only the lists of elected variables - in a SUB/BVE/BCE round no
//         kernel observes the order of any other list (the gate searches that scan foreign lists take
//         the match with the smallest clause index, elim.cu), and the table is rebuilt next round;
//   need  (one byte per literal): only the lists ERE is going to binary-search (elim.cu, k_ere_pairs).
__device__ __forceinline__ bool sortWanted(const u32* __restrict__ vinfo, const unsigned char* __restrict__ need, u32 lit) {
    if (vinfo) return (vinfo[lit >> 1] & 7u) == MIS_ELECTED;
    if (need) return need[lit] != 0;
    return true;
}
__global__ void __launch_bounds__(256) k_sort_count(const u32* __restrict__ otSize, u32 ND, DevCounters* dc, const u32* __restrict__ vinfo,
                                                    const unsigned char* __restrict__ need) {
    __shared__ u32 cnt[NCLASS];
    if (threadIdx.x < NCLASS) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (u32 lit = 2 + blockIdx.x * blockDim.x + threadIdx.x; lit < ND; lit += gridDim.x * blockDim.x) {
        const int k = sortWanted(vinfo, need, lit) ? sortClass(otSize[lit]) : -1;
        if (k >= 0) atomicAdd(&cnt[k], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NCLASS && cnt[threadIdx.x]) atomicAdd(&dc->sortCnt[threadIdx.x], cnt[threadIdx.x]);
}
// pass 2: the literals of class k go to q[start_k ...), start = exclusive sum of the class sizes.
// Each CTA owns a contiguous tile of literals, counts its classes in shared memory, reserves its
// ranges with one global atomic per class, then places its literals with shared-memory atomics.
#define FILL_TILE 4096
__global__ void __launch_bounds__(256) k_sort_fill(const u32* __restrict__ otSize, u32 ND, u32* __restrict__ q, DevCounters* dc,
                                                   const u32* __restrict__ vinfo, const unsigned char* __restrict__ need) {
    __shared__ u32 start[NCLASS], cnt[NCLASS], base[NCLASS];
    if (threadIdx.x == 0) { u32 s = 0; for (int k = 0; k < NCLASS; k++) { start[k] = s; s += dc->sortCnt[k]; } }
    for (u32 t0 = 2 + blockIdx.x * FILL_TILE; t0 < ND; t0 += gridDim.x * FILL_TILE) {
        __syncthreads();
        if (threadIdx.x < NCLASS) cnt[threadIdx.x] = 0;
        __syncthreads();
        int cls[FILL_TILE / 256]; u32 pos[FILL_TILE / 256];
#pragma unroll
        for (int k = 0; k < FILL_TILE / 256; k++) {
            const u32 lit = t0 + k * 256 + threadIdx.x;
            cls[k] = (lit < ND && sortWanted(vinfo, need, lit)) ? sortClass(otSize[lit]) : -1;
            if (cls[k] >= 0) pos[k] = atomicAdd(&cnt[cls[k]], 1u);
        }
        __syncthreads();
        if (threadIdx.x < NCLASS) base[threadIdx.x] = cnt[threadIdx.x] ? start[threadIdx.x] + atomicAdd(&dc->sortCur[threadIdx.x], cnt[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FILL_TILE / 256; k++)
            if (cls[k] >= 0) q[base[cls[k]] + pos[k]] = t0 + k * 256 + threadIdx.x;
    }
}

// ------------------------------------------------------------------ register sort (lists of <= 128)
// One group of GS lanes per list, ITEMS entries per lane (entry e lives in lane e % GS, slot
// e / GS, so list reads and writes are coalesced).  The 16-byte OLIST_CMP key is folded into
// 128 bits - size:14 | first literal:25 | last literal:25, then signature:32 | clause index:32 -
// which is exact while every clause is shorter than 2^14 literals and literals are below 2^25
// (checked on the host; otherwise the shared-memory kernels below sort everything).  Classical
// bitonic network; exchanges between lanes are shuffles, exchanges between slots stay in the
// thread.  Lists shorter than the network are padded with +inf.
// Algorithmic bytes per entry: 4 (read) + 16 (key gather) + 4 (write).
template <int GS, int ITEMS>
__global__ void __launch_bounds__(256) k_sort_reg(const uint4* __restrict__ key, const u32* __restrict__ otStart,
                                                  const u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                  const u32* __restrict__ q, const DevCounters* dc, int cls) {
    u32 qStart = 0;
    for (int k = 0; k < cls; k++) qStart += dc->sortCnt[k];
    const u32 nq = dc->sortCnt[cls];
    const u32 lane = threadIdx.x & (u32)(GS - 1);
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 rounds = (nq + groupsPerGrid - 1) / groupsPerGrid;   // every lane runs every round: shuffles are warp-wide
    u32 gi = (blockIdx.x * blockDim.x + threadIdx.x) / GS;
    for (u32 it = 0; it < rounds; it++, gi += groupsPerGrid) {
        u32 n = 0; u32* list = nullptr;
        if (gi < nq) { const u32 lit = q[qStart + gi]; n = otSize[lit]; list = occurs + otStart[lit]; }
        u64 ka[ITEMS], kb[ITEMS];
#pragma unroll
        for (int s = 0; s < ITEMS; s++) {
            const u32 e = s * GS + lane;
            ka[s] = ~0ull; kb[s] = ~0ull;
            if (e < n) {
                const u32 r = list[e];
                const uint4 k = key[r];
                ka[s] = ((u64)k.x << 50) | ((u64)k.y << 25) | (u64)k.z;
                kb[s] = ((u64)k.w << 32) | r;
            }
        }
#pragma unroll
        for (int k = 2; k <= GS * ITEMS; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= GS) {   // partner in the same lane, slot s ^ (j / GS)
#pragma unroll
                    for (int s = 0; s < ITEMS; s++) {
                        const int t = s ^ (j / GS);
                        if (t > s) {
                            const u32 e = s * GS + lane;
                            const bool asc = (e & k) == 0;
                            const bool gt = ka[s] > ka[t] || (ka[s] == ka[t] && kb[s] > kb[t]);
                            if (gt == asc) { u64 x = ka[s]; ka[s] = ka[t]; ka[t] = x; x = kb[s]; kb[s] = kb[t]; kb[t] = x; }
                        }
                    }
                } else {         // partner in lane ^ j, same slot
#pragma unroll
                    for (int s = 0; s < ITEMS; s++) {
                        const u32 e = s * GS + lane;
                        const u64 oa = __shfl_xor_sync(0xffffffffu, ka[s], j), ob = __shfl_xor_sync(0xffffffffu, kb[s], j);
                        const bool asc = (e & k) == 0;
                        const bool lower = (lane & j) == 0;           // this lane holds the lower index of the pair
                        const bool gt = ka[s] > oa || (ka[s] == oa && kb[s] > ob);   // mine > partner's
                        // the lower index keeps the smaller value when ascending
                        const bool takeOther = (lower == asc) ? gt : !gt;
                        if (takeOther) { ka[s] = oa; kb[s] = ob; }
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < ITEMS; s++) { const u32 e = s * GS + lane; if (e < n) list[e] = (u32)kb[s]; }
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

// one warp per list, up to 512 entries, shared-memory slice per warp (exact 160-bit compare)
__global__ void __launch_bounds__(MED_WARPS * 32) k_sort_med(const uint4* __restrict__ key, const u32* __restrict__ otStart,
                                                             const u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                             const u32* __restrict__ q, const DevCounters* dc, int cls0, int cls1) {
    __shared__ u64 ska[MED_WARPS][SORT_MED];
    __shared__ u64 skb[MED_WARPS][SORT_MED];
    __shared__ u32 sid[MED_WARPS][SORT_MED];
    const u32 w = threadIdx.x >> 5, l = threadIdx.x & 31u;
    u32 qStart = 0, nq = 0;   // classes cls0..cls1 are contiguous in the queue
    for (int k = 0; k < cls0; k++) qStart += dc->sortCnt[k];
    for (int k = cls0; k <= cls1; k++) nq += dc->sortCnt[k];
    for (u32 qi = blockIdx.x * MED_WARPS + w; qi < nq; qi += gridDim.x * MED_WARPS) {
        const u32 lit = q[qStart + qi], n = otSize[lit];
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
    u32 qStart = 0;
    for (int k = 0; k < 7; k++) qStart += dc->sortCnt[k];
    const u32 nq = dc->sortCnt[7];
    for (u32 qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        const u32 lit = q[qStart + qi], n = otSize[lit];
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
    u32 qStart = 0;
    for (int k = 0; k < 8; k++) qStart += dc->sortCnt[k];
    const u32 nq = dc->sortCnt[8];
    for (u32 qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        const u32 lit = q[qStart + qi], n = otSize[lit];
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

__global__ void k_sort_reset(DevCounters* dc) {
    for (int k = 0; k < NCLASS; k++) { dc->sortCnt[k] = 0; dc->sortCur[k] = 0; }
}

// mode 0: every list, 1: lists of elected variables, 2: lists flagged in needSort
void launchSortOT(Ctx* c, int mode) {
    const u32* vinfo = mode == 1 ? c->rank : nullptr;
    const unsigned char* need = mode == 2 ? c->needSort : nullptr;
    const size_t bigSmem = (size_t)SORT_BIG * (8 + 8 + 4);
    if (!c->attrSort) {
        cudaFuncSetAttribute(k_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bigSmem);
        c->attrSort = true;
    }
    u32* q = c->qMed;   // ND entries: the literals with >= 2 occurrences, grouped by length class
    LAUNCH(c, k_sort_reset, 1, 1, 0, c->dc);
    LAUNCH(c, k_sort_count, gridFor(c->ND, 256, 4), 256, 0, c->otSize, c->ND, c->dc, vinfo, need);
    LAUNCH(c, k_sort_fill, gridFor(c->ND, 256, FILL_TILE / 256), 256, 0, c->otSize, c->ND, q, c->dc, vinfo, need);
    // the folded 128-bit key is exact while literals < 2^25 and clauses are shorter than 2^14 (flag 8: k_ot_count)
    const bool fold = c->ND <= (1u << 25) && !(c->hdc->flags & 8u);
    // grids: enough groups for every list of a class if all of them fell into it, capped
    const u32 nLists = c->ND;
    auto grid = [&](u32 gs) { u64 b = ((u64)nLists * gs + 255) / 256; return (u32)(b > 148ull * 16 ? 148ull * 16 : (b ? b : 1)); };
    if (fold) {
        LAUNCH(c, (k_sort_reg<4, 1>), grid(4), 256, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 0);
        LAUNCH(c, (k_sort_reg<8, 1>), grid(8), 256, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 1);
        LAUNCH(c, (k_sort_reg<16, 1>), grid(16), 256, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 2);
        LAUNCH(c, (k_sort_reg<32, 1>), grid(32), 256, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 3);
        LAUNCH(c, (k_sort_reg<32, 2>), grid(32), 256, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 4);
        LAUNCH(c, (k_sort_reg<32, 4>), grid(32), 256, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 5);
        LAUNCH(c, k_sort_med, 148 * 4, MED_WARPS * 32, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 6, 6);
    } else
        LAUNCH(c, k_sort_med, 148 * 4, MED_WARPS * 32, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc, 0, 6);
    LAUNCH(c, k_sort_big, 148, 512, bigSmem, c->key, c->otStart, c->otSize, c->occurs, q, c->dc);
    LAUNCH(c, k_sort_huge, 32, 1024, 0, c->key, c->otStart, c->otSize, c->occurs, q, c->dc);
}
