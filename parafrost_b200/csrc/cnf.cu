// parafrost_b200/csrc/cnf.cu -- clause store kernels: awaken/prep (+ the first round's histogram), literal
// histogram + sort keys, occurrence-table construction (radix partition + staged placement), live counts,
// garbage-collecting compaction, result store (+ the -aggresivesort write-back order).
//
// Reference behaviour being replaced (results must be identical):
//   prep_cnf_k            src/gpu/cnf.cu:45-53        sort literals, 32-bit signature
//   copy_if_k + histSimp  src/gpu/cnf.cu:33-43, histogram.cu:54-72  (thrust sort only to count!)
//   create_ot_k           src/gpu/occurrence.cu:50-62   (k_ot_part + k_ot_place)
//   cnt_cls_lits          src/gpu/count.cu:83-106
//   scatter_k/compact_k   src/gpu/recycle.cu:34-105
//   cacheCNF              src/gpu/cnf.cu:200-237
//   thrust::stable_sort   src/gpu/cnf.cu:232-233      (-aggresivesort: aggressiveOrder)
// All of them stream the clause store once, coalesced; what bounds each is measured in DESIGN.md 7.
#include <cstdlib>

#include "common.cuh"

// ------------------------------------------------------------------ awaken + prep
// algorithmic bytes: read 8(C+1) offsets + 4L literals [+4C meta], write 16C headers + 4L literals
// With hist != NULL the first round's histogram + sort-key pass (k_hist_key) is fused in: the sorted
// literals are in registers anyway, so the clause store is not read again before the partition.
__global__ void k_awaken(const u32* __restrict__ inLits, const u64* __restrict__ inOffs, const u32* __restrict__ inMeta,
                         u64 C, uint4* __restrict__ hdr, u32* __restrict__ pool, u32* __restrict__ hist, uint4* __restrict__ key, u32* flags,
                         u32 ND, u64 L0) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < C; i += (u64)gridDim.x * blockDim.x) {
        const u64 b = inOffs[i];
        u64 e = inOffs[i + 1];
        // input validation (the C ABI takes raw buffers): offsets must be monotone and inside the literal array, literals
        // inside [2, 2V+2).  A bad entry is neutralised (empty clause / literal 2) and flagged: the call fails with
        // SIGMA_BAD_ARGUMENT at its first read-back instead of indexing outside the tables.
        if (e < b || e > L0 || e - b >= (1ull << 31)) { atomicOr(flags, 128u); e = b; }
        const int sz = (int)(e - b);
        u32* dst = pool + b;
        u32 sig = 0;
        if (sz <= 8) {
            u32 r[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                r[k] = (k < sz) ? inLits[b + k] : 0xFFFFFFFFu;
                if (k < sz && (r[k] < 2u || r[k] >= ND)) { atomicOr(flags, 128u); r[k] = 2u; }
            }
            // odd-even transposition network on 8 registers (padding sorts to the end)
#pragma unroll
            for (int pass = 0; pass < 8; pass++) {
#pragma unroll
                for (int k = (pass & 1); k + 1 < 8; k += 2) {
                    const u32 a = r[k], bb = r[k + 1];
                    r[k] = min(a, bb); r[k + 1] = max(a, bb);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < sz) { dst[k] = r[k]; sig |= MAPHASH(r[k]); }
        } else {
            for (int k = 0; k < sz; k++) {
                u32 t = inLits[b + k];
                if (t < 2u || t >= ND) { atomicOr(flags, 128u); t = 2u; }
                int j = k;
                for (; j > 0 && t < dst[j - 1]; j--) dst[j] = dst[j - 1];
                dst[j] = t;
                sig |= MAPHASH(t);
            }
        }
        if (sz <= 1) sig = 0;  // calcSig leaves the signature untouched for size <= 1 (primitives.cuh:177-185)
        u32 bits = 0;
        if (inMeta) {
            const u32 m = inMeta[i];
            if (m & CB_LEARNT) bits = m & ~(CB_DELETED | CB_MOLTEN | CB_ADDED);
        }
        hdr[i] = make_uint4((u32)b, (u32)sz, sig, bits);
        if (hist) {
            for (int k = 0; k < sz; k++) atomicAdd(&hist[dst[k]], 1u);
            key[i] = make_uint4((u32)sz, sz ? dst[0] : 0u, sz ? dst[sz - 1] : 0u, sig);
            if (sz >= (1 << 14)) atomicOr(flags, 8u);
        }
    }
}

static void launchCountPass(Ctx* c, bool awaken, u32 n, u64 numLiterals, u64 numClauses);
static void launchScatter2(Ctx* c, u32 n);
static bool otV2();

void launchAwaken(Ctx* c) {
    if (!c->C0) return;
    if (otV2()) { launchCountPass(c, true, (u32)c->C0, c->L0, c->C0); c->histFresh = true; return; }
    cudaMemsetAsync(c->hist, 0, (size_t)c->ND * 4, c->stream);
    LAUNCH(c, k_awaken, gridFor(c->C0, 256), 256, 0, c->inLits, c->inOffs, c->inMeta, c->C0, c->hdr[c->cur], c->pool[c->cur], c->hist, c->key,
           &c->dc->flags, c->ND, c->L0);
    KB(c, 8.0 * c->C0 + 4.0 * c->L0 + (c->inMeta ? 4.0 * c->C0 : 0.0) + 16.0 * c->C0 + 4.0 * c->L0 + 16.0 * c->C0 + 4.0 * c->ND);   // offsets + literals in, headers + literals + keys + histogram out
    c->histFresh = true;   // hist[] and key[] describe the store until a kernel changes it (api.cu: buildOT)
}

// ------------------------------------------------------------------ histogram + sort keys
// algorithmic bytes: read 16C + 4L, 4 per literal of atomic traffic on hist (L2 resident), write 16C keys
__global__ void k_hist_key(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n,
                           u32* __restrict__ hist, uint4* __restrict__ key, u32* flags) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        const u32* l = pool + h.x;
        const int sz = (int)h.y;
        u32 first = 0, last = 0;
        for (int k = 0; k < sz; k++) {
            const u32 lit = l[k];
            if (k == 0) first = lit;
            last = lit;
            atomicAdd(&hist[lit], 1u);
        }
        key[i] = make_uint4(h.y, first, last, h.z);
        if (sz >= (1 << 14)) atomicOr(flags, 8u);   // the list sort's folded key needs size < 2^14 (otsort.cu)
    }
}

void launchHistKey(Ctx* c) {
    if (otV2()) { if (c->hdc->numCls) launchCountPass(c, false, c->hdc->numCls, c->numLiterals, c->numClauses); return; }
    cudaMemsetAsync(c->hist, 0, (size_t)c->ND * 4, c->stream);
    const u32 n = c->hdc->numCls;
    if (n) {
        LAUNCH(c, k_hist_key, gridFor(n, 256), 256, 0, c->hdr[c->cur], c->pool[c->cur], n, c->hist, c->key, &c->dc->flags);
        KB(c, 16.0 * n + 4.0 * c->numLiterals + 16.0 * c->numClauses + 4.0 * c->ND);
    }
}

// ------------------------------------------------------------------ occurrence lists: partition + place
// create_ot_k (occurrence.cu:50-62) appends every clause reference to the lists of its literals
// with one global atomic and one random 8-byte store per literal.  Random 4-byte stores over an
// occurs[] array far larger than L2 cost a DRAM sector each, so the lists are built in two
// streaming passes instead (an MSD radix partition by literal, the histogram being known):
//   k_ot_part   streams the clause store once; every CTA bins the (literal, clause) pairs of its
//               tile by literal range ("bucket" = 2^shift consecutive literals, <= 1024 buckets for
//               V <= 2^24), reserves one run per bucket with a single global atomic and writes the
//               pairs into the bucket's segment of pairs[] (the segment bounds are otStart[] at the
//               bucket borders: no extra histogram).  Runs of neighbouring CTAs complete each
//               other's sectors in L2.
//   k_ot_place  one CTA per bucket: list cursors of the bucket's literals in shared memory, pairs
//               streamed once (coalesced), clause indices stored into occurs[] inside the bucket's
//               window (<= a few hundred KB: the sectors are completed in L2 before they reach HBM).
// The order inside a list is arbitrary (as in the reference); k_sort_* fixes it afterwards.
// Measured (profiles/r01_ncu_full_cfg2_v3.txt): both kernels are bound by the LSU/MIO path - one
// shared-memory atomic (~2 cycles per lane) and one 4/8-byte store to its own sector per pair - not
// by HBM.  A warp-match ranking variant (ballots + warp-private counters, no atomics) was measured
// 10-100 % slower on cfg2-cfg4 and dropped; clause-local formulas (Tseitin, multiplier) already
// write long runs per bucket and partition at ~3x the speed of uniform random k-SAT.
// Algorithmic bytes: part 16C + 4L read + 8L written; place 8L read + 4L written + 12 ND.
#define PART_THREADS 1024
#define PART_SHORT 8       // clauses up to this size keep the ranks of their literals in registers
#define PART_STAGE 16384u  // pairs of a tile staged in shared memory (128 KB)

// One shared-memory atomic per pair: the rank the counting sweep hands out IS the pair's slot in the
// tile's run of its bucket, so it is kept (16 bits per literal, four registers per clause) and the
// writing sweep needs no second atomic.  Literals of longer clauses are counted separately and take
// the tail of the run with a second atomic.
// The writing sweep goes through shared memory: the tile's pairs are laid out bucket by bucket
// (tileOff = exclusive scan of the tile's bucket counts) and copied out in that order, so that
// neighbouring lanes store to neighbouring addresses of a run.  Per-lane scattered 8-byte stores cost
// one L2 write transaction each, and that transaction rate - not HBM - bounded the unstaged kernel.
// A tile with more pairs than the stage holds (long clauses) writes directly.
// PART_CPT clauses per thread: 3 (tiles of 3072 clauses) or 5 for short clauses, so that a tile fills the stage;
// the first PART_KEEP literals of a clause stay in registers between the two sweeps (the 220 KB of shared
// memory leave almost no L1, a second read would come from L2)
template <int PART_CPT, int PART_KEEP>
__global__ void __launch_bounds__(PART_THREADS) k_ot_part(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n,
                                                          const u32* __restrict__ otStart, u32 ND, u32 shift, u32 NB, u32 stageCap,
                                                          u32* __restrict__ gcur, uint2* __restrict__ pairs) {
    extern __shared__ u32 sm[];
    u32* cntS = sm;             // literals of short clauses per bucket
    u32* cntL = sm + NB;        // literals of long clauses per bucket, then their running slot
    u32* gbase = sm + 2 * NB;   // start of this tile's run in the bucket's segment
    u32* tileOff = sm + 3 * NB; // start of the bucket inside the staged tile
    uint2* stage = (uint2*)(sm + 4 * NB);   // 16 NB bytes: 8-byte aligned
    __shared__ u32 warpTot[32];
    __shared__ u32 tileTotal, nonEmpty;
    const u32 tile0 = blockIdx.x * (PART_THREADS * PART_CPT);
    for (u32 b = threadIdx.x; b < 2 * NB; b += PART_THREADS) sm[b] = 0;
    if (threadIdx.x == 0) nonEmpty = 0;
    u32 off[PART_CPT], sz[PART_CPT], rk[PART_CPT][PART_SHORT / 2], lk[PART_CPT][PART_KEEP];
#pragma unroll
    for (int k = 0; k < PART_CPT; k++) {
        const u32 i = tile0 + k * PART_THREADS + threadIdx.x;
        sz[k] = 0; off[k] = 0;
        if (i < n) {
            const uint4 h = hdr[i];
            if (!C_DELETED(h.w)) { off[k] = h.x; sz[k] = h.y; }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PART_CPT; k++) {
        const u32* l = pool + off[k];
#pragma unroll
        for (int q = 0; q < PART_SHORT / 2; q++) rk[k][q] = 0;
        if (sz[k] <= PART_SHORT) {
#pragma unroll
            for (int q = 0; q < PART_SHORT; q++)
                if ((u32)q < sz[k]) {
                    const u32 lit = l[q];
                    if (q < PART_KEEP) lk[k][q < PART_KEEP ? q : 0] = lit;
                    rk[k][q >> 1] |= atomicAdd(&cntS[lit >> shift], 1u) << ((q & 1) * 16);
                }
        } else
            for (u32 q = 0; q < sz[k]; q++) atomicAdd(&cntL[l[q] >> shift], 1u);
    }
    __syncthreads();
    // per bucket: reserve the run (one global atomic), exclusive scan of the tile's bucket counts
    const u32 per = (NB + PART_THREADS - 1) / PART_THREADS;   // consecutive buckets per thread
    const u32 b0 = threadIdx.x * per;
    u32 mine = 0, used = 0;
    for (u32 q = 0; q < per; q++) {
        const u32 b = b0 + q;
        if (b < NB) {
            const u32 tS = cntS[b], tL = cntL[b];
            if (tS + tL) { gbase[b] = otStart[min(b << shift, ND)] + atomicAdd(&gcur[b], tS + tL); used++; }
            cntL[b] = tS;   // long-clause literals follow the short ones
            tileOff[b] = mine;
            mine += tS + tL;
        }
    }
    const u32 incl = warpIncl(mine);
    used = warpSum(used);
    if ((threadIdx.x & 31u) == 0 && used) atomicAdd(&nonEmpty, used);
    if ((threadIdx.x & 31u) == 31u) warpTot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const u32 t = warpTot[threadIdx.x];
        const u32 ti = warpIncl(t);
        warpTot[threadIdx.x] = ti - t;
        if (threadIdx.x == 31) tileTotal = ti;
    }
    __syncthreads();
    const u32 base = warpTot[threadIdx.x >> 5] + incl - mine;
    for (u32 q = 0; q < per; q++) { const u32 b = b0 + q; if (b < NB) tileOff[b] += base; }
    __syncthreads();
    // Staging pays when the tile's runs are short (uniform random formulas: ~4 pairs per bucket).  In
    // clause-local formulas (Tseitin, arithmetic) neighbouring clauses fall into the same bucket with
    // consecutive ranks, so the direct stores of a warp already coalesce and staging only adds a pass.
    const bool staged = tileTotal <= stageCap && tileTotal < 12u * nonEmpty;
#pragma unroll
    for (int k = 0; k < PART_CPT; k++) {
        const u32 i = tile0 + k * PART_THREADS + threadIdx.x;
        const u32* l = pool + off[k];
        if (sz[k] <= PART_SHORT) {
#pragma unroll
            for (int q = 0; q < PART_SHORT; q++)
                if ((u32)q < sz[k]) {
                    const u32 lit = q < PART_KEEP ? lk[k][q < PART_KEEP ? q : 0] : l[q];
                    const u32 b = lit >> shift;
                    const u32 r = (rk[k][q >> 1] >> ((q & 1) * 16)) & 0xFFFFu;
                    if (staged) stage[tileOff[b] + r] = make_uint2(lit, i);
                    else pairs[gbase[b] + r] = make_uint2(lit, i);
                }
        } else
            for (u32 q = 0; q < sz[k]; q++) {
                const u32 lit = l[q];
                const u32 b = lit >> shift;
                const u32 r = atomicAdd(&cntL[b], 1u);
                if (staged) stage[tileOff[b] + r] = make_uint2(lit, i);
                else pairs[gbase[b] + r] = make_uint2(lit, i);
            }
    }
    if (!staged) return;
    __syncthreads();
    const u32 total = tileTotal;
    for (u32 t = threadIdx.x; t < total; t += PART_THREADS) {
        const uint2 pr = stage[t];
        const u32 b = pr.x >> shift;
        pairs[gbase[b] + (t - tileOff[b])] = pr;
    }
}

#define PLACE_THREADS 1024
#define PLACE_SPLIT 64           // a bucket too large for the window is shared by up to 64 work units ...
#define PLACE_UNIT (32u << 10)   // ... of about this many pairs each
#define PLACE_WINDOW (40u << 10) // entries of a bucket's occurs[] window that can be staged in shared memory
// Staged mode (the normal case, buckets are sized for it): the whole occurs[] window of the bucket is
// assembled in shared memory - one shared-memory atomic and one shared-memory store per pair - and
// written out with coalesced full-sector stores.  Scattering the 4-byte entries straight into global
// memory costs one L2 write transaction per entry, and ~65 G transactions/s chip-wide was the
// measured ceiling of the unstaged kernel (profiles/r01_ncu_full_cfg2_v3.txt).
// A bucket too large for the window (hot literal ranges of structured formulas: buckets are literal
// RANGES, so their pair counts follow the formula's structure - multiplier inputs: one bucket with
// 28x the mean) is queued as work units for k_ot_place_big.
__global__ void __launch_bounds__(PLACE_THREADS) k_ot_place(const uint2* __restrict__ pairs, const u32* __restrict__ otStart, u32 ND,
                                                            u32 shift, u32 window, u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                            u32* __restrict__ big, u32* nBig) {
    extern __shared__ u32 smem[];
    const u32 W = 1u << shift;
    u32* cur = smem;          // [W] list cursors
    u32* win = smem + W;      // [window] staged entries
    const u32 lit0 = blockIdx.x << shift;
    const u32 litEnd = min(lit0 + W, ND);
    const u32 p0 = otStart[lit0], p1 = otStart[litEnd];
    const u32 len = p1 - p0;
    if (len > window) {
        if (threadIdx.x == 0) {
            u32 S = (len + PLACE_UNIT - 1) / PLACE_UNIT;
            S = S > PLACE_SPLIT ? PLACE_SPLIT : S;
            const u32 base = atomicAdd(nBig, S);
            for (u32 s = 0; s < S; s++) big[base + s] = blockIdx.x | (s << 13) | (S << 19);   // bucket < 2^13, s < 2^6, S <= 2^6
        }
        for (u32 k = threadIdx.x; lit0 + k < litEnd; k += PLACE_THREADS) otSize[lit0 + k] = 0;   // global list cursors of the work units
        return;
    }
    const u32 nl = litEnd - lit0;
    for (u32 k = threadIdx.x; k < nl; k += PLACE_THREADS) cur[k] = otStart[lit0 + k] - p0;
    __syncthreads();
    u32 j = p0 + threadIdx.x;
    for (; j + 3 * PLACE_THREADS < p1; j += 4 * PLACE_THREADS) {
        uint2 p[4];
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = pairs[j + k * PLACE_THREADS];
#pragma unroll
        for (int k = 0; k < 4; k++) win[atomicAdd(&cur[p[k].x - lit0], 1u)] = p[k].y;
    }
    for (; j < p1; j += PLACE_THREADS) { const uint2 p = pairs[j]; win[atomicAdd(&cur[p.x - lit0], 1u)] = p.y; }
    __syncthreads();
    for (u32 k = threadIdx.x; k < len; k += PLACE_THREADS) occurs[p0 + k] = win[k];
    for (u32 k = threadIdx.x; k < nl; k += PLACE_THREADS) otSize[lit0 + k] = cur[k] - (otStart[lit0 + k] - p0);
}
// Work unit (b, s of S): the s-th share of the PAIRS of an oversized bucket.  The list cursors of such a
// bucket live in global memory (otSize[], zeroed by k_ot_place; L2 resident: the bucket is one
// literal range), so every pair is read once and placed with one L2 atomic.
__global__ void __launch_bounds__(PLACE_THREADS) k_ot_place_big(const uint2* __restrict__ pairs, const u32* __restrict__ otStart, u32 ND,
                                                                u32 shift, u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                                const u32* __restrict__ big, const u32* nBig) {
    const u32 W = 1u << shift;
    const u32 nItems = *nBig;
    for (u32 item = blockIdx.x; item < nItems; item += gridDim.x) {
        const u32 code = big[item];
        const u32 b = code & 0x1FFFu, s = (code >> 13) & 0x3Fu, S = code >> 19;
        const u32 lit0 = b << shift;
        const u32 p0 = otStart[lit0], p1 = otStart[min(lit0 + W, ND)];
        const u32 len = p1 - p0;
        const u32 q0 = p0 + (u32)((u64)len * s / S), q1 = p0 + (u32)((u64)len * (s + 1) / S);
        u32 j = q0 + threadIdx.x;
        for (; j + 3 * PLACE_THREADS < q1; j += 4 * PLACE_THREADS) {
            uint2 p[4]; u32 pos[4];
#pragma unroll
            for (int k = 0; k < 4; k++) p[k] = pairs[j + k * PLACE_THREADS];
#pragma unroll
            for (int k = 0; k < 4; k++) pos[k] = otStart[p[k].x] + atomicAdd(&otSize[p[k].x], 1u);
#pragma unroll
            for (int k = 0; k < 4; k++) occurs[pos[k]] = p[k].y;
        }
        for (; j < q1; j += PLACE_THREADS) {
            const uint2 p = pairs[j];
            occurs[otStart[p.x] + atomicAdd(&otSize[p.x], 1u)] = p.y;
        }
    }
}

// ------------------------------------------------------------------ placement with the bulk-copy engine (sm_90+ / sm_100a)
// k_ot_place is bound by the latency of its pair loads (ncu: 22 of 40 stall cycles per issue are long-scoreboard) and ends
// with 1024 threads copying a contiguous window out of shared memory.  Both are what the TMA unit is for:
//   * the bucket's pairs - one contiguous, 16-byte-alignable range of otPairs[] - arrive through `cp.async.bulk` into a
//     two-stage shared-memory ring (16 KB per stage), completion signalled on an mbarrier; the 1024 threads only ever
//     read shared memory, the next chunk is in flight while the current one is placed;
//   * the finished window leaves as ONE `cp.async.bulk` shared -> global store (the window is laid out with the same
//     16-byte phase as its destination; the unaligned head and tail - at most 3 entries each - are stored by threads).
// Same result as k_ot_place (list order inside a bucket is arbitrary in both: shared-memory atomics hand out the slots).
// Built and measured in round 2, NOT the default: 15 % slower than k_ot_place (see launchScatter2); SIGMA_OT_TMA=1 selects it.
#define PLACE_CH 2048u   // pairs per ring stage
__device__ __forceinline__ u32 smemU32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count)); }
__device__ __forceinline__ void mbarExpectTx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemU32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(u64* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smemU32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gsrc, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemU32(smemDst)), "l"(gsrc), "r"(bytes), "r"(smemU32(bar)) : "memory");
}
__device__ __forceinline__ void bulkStore(void* gdst, const void* smemSrc, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smemU32(smemSrc)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(PLACE_THREADS) k_ot_place_tma(const uint2* __restrict__ pairs, const u32* __restrict__ otStart, u32 ND,
                                                                u32 shift, u32 window, u32* __restrict__ otSize, u32* __restrict__ occurs,
                                                                u32* __restrict__ big, u32* nBig) {
    extern __shared__ __align__(128) u32 smem[];
    const u32 W = 1u << shift;
    u32* cur = smem;                                      // [W] list cursors
    u32* win = smem + W;                                  // [window + 8] staged entries, same 16-byte phase as occurs + p0
    uint2* ring = (uint2*)(smem + W + window + 8);        // [2][PLACE_CH]
    u64* bar = (u64*)(ring + 2 * PLACE_CH);               // [2]
    const u32 lit0 = blockIdx.x << shift;
    const u32 litEnd = min(lit0 + W, ND);
    const u32 p0 = otStart[lit0], p1 = otStart[litEnd];
    const u32 len = p1 - p0;
    if (len > window) {
        if (threadIdx.x == 0) {
            u32 S = (len + PLACE_UNIT - 1) / PLACE_UNIT;
            S = S > PLACE_SPLIT ? PLACE_SPLIT : S;
            const u32 base = atomicAdd(nBig, S);
            for (u32 s = 0; s < S; s++) big[base + s] = blockIdx.x | (s << 13) | (S << 19);
        }
        for (u32 k = threadIdx.x; lit0 + k < litEnd; k += PLACE_THREADS) otSize[lit0 + k] = 0;
        return;
    }
    const u32 nl = litEnd - lit0;
    const u32 a = p0 & 3u;                                // entry k of the window sits at win[a + k]
    for (u32 k = threadIdx.x; k < nl; k += PLACE_THREADS) cur[k] = otStart[lit0 + k] - p0 + a;
    // the pair range, widened to 16-byte alignment (a pair is 8 bytes): [q0, q1) contains [p0, p1)
    const u32 q0 = p0 & ~1u, q1 = (p1 + 1u) & ~1u;
    const u32 nCh = (q1 - q0 + PLACE_CH - 1) / PLACE_CH;
    if (threadIdx.x == 0) {
        mbarInit(&bar[0], 1); mbarInit(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (u32 c = 0; c < 2 && c < nCh; c++) {
            const u32 n = min(PLACE_CH, q1 - (q0 + c * PLACE_CH));
            mbarExpectTx(&bar[c], n * 8u);
            bulkLoad(ring + c * PLACE_CH, pairs + q0 + c * PLACE_CH, n * 8u, &bar[c]);
        }
    for (u32 c = 0; c < nCh; c++) {
        const u32 st = c & 1u;
        mbarWait(&bar[st], (c >> 1) & 1u);
        const u32 base = q0 + c * PLACE_CH;
        const u32 n = min(PLACE_CH, q1 - base);
        const uint2* src = ring + st * PLACE_CH;
        for (u32 j = threadIdx.x; j < n; j += PLACE_THREADS) {
            const u32 gidx = base + j;
            if (gidx >= p0 && gidx < p1) { const uint2 p = src[j]; win[atomicAdd(&cur[p.x - lit0], 1u)] = p.y; }
        }
        __syncthreads();   // every reader of this stage is done: it may be refilled
        if (threadIdx.x == 0 && c + 2 < nCh) {
            const u32 n2 = min(PLACE_CH, q1 - (base + 2 * PLACE_CH));
            mbarExpectTx(&bar[st], n2 * 8u);
            bulkLoad(ring + st * PLACE_CH, pairs + base + 2 * PLACE_CH, n2 * 8u, &bar[st]);
        }
    }
    // window -> occurs[p0 .. p1): the 16-byte-aligned middle as one bulk store, head and tail by threads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of win[] -> visible to the bulk engine
    __syncthreads();
    const u32 k0 = (4u - a) & 3u;                          // first entry whose global index is a multiple of 4
    const u32 mid = len > k0 ? ((len - k0) & ~3u) : 0u;
    if (threadIdx.x == 0 && mid) {
        for (u32 o = 0; o < mid; o += 8192u) bulkStore(occurs + p0 + k0 + o, win + a + k0 + o, min(8192u, mid - o) * 4u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (threadIdx.x < k0 && threadIdx.x < len) occurs[p0 + threadIdx.x] = win[a + threadIdx.x];            // head: < 4 entries
    for (u32 k = k0 + mid + threadIdx.x; k < len; k += PLACE_THREADS) occurs[p0 + k] = win[a + k];       // tail: < 4 entries (or everything when mid == 0)
    for (u32 k = threadIdx.x; k < nl; k += PLACE_THREADS) otSize[lit0 + k] = cur[k] - (otStart[lit0 + k] - p0 + a);
    if (threadIdx.x == 0 && mid) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

void launchScatter(Ctx* c) {
    const u32 n = c->hdc->numCls;
    // nothing live (e.g. prop() satisfied every clause): empty lists; hist/otStart may be stale here
    if (!n || !c->numLiterals) {
        cudaMemsetAsync(c->otSize, 0, (size_t)c->ND * 4, c->stream);
        if (otV2()) cudaMemsetAsync(c->hist, 0, (size_t)c->ND * 4, c->stream);
        return;
    }
    if (otV2()) { launchScatter2(c, n); return; }
    if (!c->attrOT) {
        cudaFuncSetAttribute(k_ot_part<3, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_part<5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_place, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (PLACE_WINDOW + (1 << 12)));
        c->attrOT = true;
    }
    // bucket = 2^shift consecutive literals, sized so that an average bucket fills at most ~70 % of a
    // staging window; at most 8192 buckets (shared-memory counters of k_ot_part)
    u32 shift = 6;
    while (shift < 12 && ((u64)c->numLiterals << (shift + 1)) / c->ND <= PLACE_WINDOW * 7 / 10) shift++;
    while (shift < 15 && ((c->ND + (1u << shift) - 1) >> shift) > 8192) shift++;
    const u32 NB = (c->ND + (1u << shift) - 1) >> shift;
    c->otShift = shift; c->otNB = NB;
    // beyond 2^12 literals per bucket (more than 2^25 literals in all) the cursors alone fill the shared memory: direct mode only
    u32 window = shift <= 12 ? PLACE_WINDOW : 0;
    if (const char* w = getenv("SIGMA_OT_WINDOW")) { const u32 v = (u32)atoi(w); if (v < window) window = v; }   // tests: force the work-unit path
    const size_t placeSmem = shift <= 12 ? 4 * ((size_t)PLACE_WINDOW + (1u << shift)) : (size_t)4 << shift;
    cudaMemsetAsync(c->otCur, 0, (size_t)NB * 4, c->stream);
    // shared memory of k_ot_part: 4 words per bucket + the stage (whatever is left of ~220 KB, at most PART_STAGE pairs)
    const size_t partFixed = 16 * (size_t)NB + 8;
    u32 stageCap = partFixed + 8 * (size_t)PART_STAGE <= 220 * 1024 ? PART_STAGE : (u32)((220 * 1024 - partFixed) / 8);
    if (c->numLiterals <= 3 * c->numClauses)   // short clauses: more of them per tile
        LAUNCH(c, (k_ot_part<5, 3>), divup(n, PART_THREADS * 5), PART_THREADS, partFixed + 8 * (size_t)stageCap, c->hdr[c->cur], c->pool[c->cur], n,
               c->otStart, c->ND, shift, NB, stageCap, c->otCur, c->otPairs);
    else
        LAUNCH(c, (k_ot_part<3, 5>), divup(n, PART_THREADS * 3), PART_THREADS, partFixed + 8 * (size_t)stageCap, c->hdr[c->cur], c->pool[c->cur], n,
               c->otStart, c->ND, shift, NB, stageCap, c->otCur, c->otPairs);
    KB(c, 16.0 * n + 4.0 * c->numLiterals + 8.0 * c->numLiterals);   // headers + literals in, (literal, clause) pairs out
    u32* nBig = &c->dc->scratch[7];
    cudaMemsetAsync(nBig, 0, 4, c->stream);
    LAUNCH(c, k_ot_place, NB, PLACE_THREADS, placeSmem, c->otPairs, c->otStart, c->ND, shift, window, c->otSize, c->occurs, c->otBig, nBig);
    KB(c, 8.0 * c->numLiterals + 4.0 * c->numLiterals + 12.0 * c->ND);   // pairs in, list entries out, list bounds
    LAUNCH(c, k_ot_place_big, 148 * 2, PLACE_THREADS, 0, c->otPairs, c->otStart, c->ND, shift, c->otSize, c->occurs, c->otBig, nBig);
}

// ================================================================== occurrence-table build, version 2
// Three passes over ~L items each used to pay one atomic per item: the per-literal histogram (global RED), the
// partition's rank (shared-memory atomic) and the placement's cursor (shared-memory atomic); the RED pass and the
// shared-memory atomics - not HBM - bounded all three (profiles/r02_bench_cfg2_v2.json: k_awaken 0.33, k_ot_part 0.25 of
// the copy roofline).  Version 2 pays two:
//   k_ot_count  (fused into awaken / the key pass, tile = 1024 x CPT clauses per CTA): one shared-memory atomic per literal on
//               the tile's BUCKET counter; the value it returns is the literal's rank inside (tile, bucket) and is kept -
//               8 x 16 bits per clause, one coalesced 16-byte store - together with the tile's row of bucket counts.
//               No global histogram, no global reductions.
//   k_ot_colsum / k_ot_bscan: bucket totals = column sums of the count matrix, exclusive scan -> segment starts.
//   k_ot_part2  the same tile again: its row of counts gives run lengths (one global atomic per non-empty (tile, bucket)
//               reserves the run - issued as independent atomics, one round trip), the stored ranks give every pair its
//               slot: NO shared-memory atomics, no counter clearing; pairs are staged bucket by bucket and copied out
//               with one table look-up per pair.  Literals of clauses longer than 8 take the tail of the run through
//               a second counter (rare).
//   k_ot_place2 one CTA per bucket as before; the per-literal histogram of the bucket (which IS hist[] / otSize[], and
//               whose local scan is otStart[]) is counted here with the one shared-memory atomic the placement needs
//               anyway: sweep 1 counts and keeps each pair's rank in registers, sweep 2 re-reads the pairs from L2 and
//               stores entry = start[literal] + rank into the staged window.
// Oversized buckets (structured formulas) keep the global-atomic work units: count, scan, place (k_ot_big_*).
#define OT_T 1024
#define OT_KEEP 5

__device__ __forceinline__ void sort8(u32 (&r)[8]) {
#pragma unroll
    for (int pass = 0; pass < 8; pass++) {
#pragma unroll
        for (int k = (pass & 1); k + 1 < 8; k += 2) {
            const u32 a = r[k], bb = r[k + 1];
            r[k] = min(a, bb); r[k + 1] = max(a, bb);
        }
    }
}

template <int CPT, bool AWAKEN>
__global__ void __launch_bounds__(OT_T) k_ot_count(const u32* __restrict__ inLits, const u64* __restrict__ inOffs, const u32* __restrict__ inMeta, u64 L0,
                                                   uint4* __restrict__ hdr, u32* __restrict__ pool, u32 n, u32 ND, u32 shift, u32 NB, u32 NBp,
                                                   uint4* __restrict__ rk8, u32* __restrict__ cntMat, uint4* __restrict__ key, u32* flags) {
    extern __shared__ u32 sm[];
    u32* cntS = sm;        // literals of clauses with <= 8 literals: their ranks are kept
    u32* cntL = sm + NB;   // literals of longer clauses: counted only
    for (u32 b = threadIdx.x; b < 2 * NB; b += OT_T) sm[b] = 0;
    __syncthreads();
    const u32 tile0 = blockIdx.x * (OT_T * CPT);
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const u32 i = tile0 + k * OT_T + threadIdx.x;
        if (i >= n) continue;
        u32 r[8];
        int sz; u32 off, sig;
        if (AWAKEN) {
            const u64 b = inOffs[i];
            u64 e = inOffs[i + 1];
            // input validation (the C ABI takes raw buffers): a bad entry is neutralised and flagged, the call fails with
            // SIGMA_BAD_ARGUMENT at its first read-back instead of indexing outside the tables
            if (e < b || e > L0 || e - b >= (1ull << 31)) { atomicOr(flags, 128u); e = b; }
            sz = (int)(e - b); off = (u32)b; sig = 0;
            u32* dst = pool + b;
            if (sz <= 8) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    r[q] = (q < sz) ? inLits[b + q] : 0xFFFFFFFFu;
                    if (q < sz && (r[q] < 2u || r[q] >= ND)) { atomicOr(flags, 128u); r[q] = 2u; }
                }
                sort8(r);
#pragma unroll
                for (int q = 0; q < 8; q++) if (q < sz) { dst[q] = r[q]; sig |= MAPHASH(r[q]); }
            } else {
                for (int q = 0; q < sz; q++) {
                    u32 t = inLits[b + q];
                    if (t < 2u || t >= ND) { atomicOr(flags, 128u); t = 2u; }
                    int j = q;
                    for (; j > 0 && t < dst[j - 1]; j--) dst[j] = dst[j - 1];
                    dst[j] = t;
                    sig |= MAPHASH(t);
                }
            }
            if (sz <= 1) sig = 0;  // calcSig leaves the signature untouched for size <= 1 (primitives.cuh:177-185)
            u32 bits = 0;
            if (inMeta) { const u32 m = inMeta[i]; if (m & CB_LEARNT) bits = m & ~(CB_DELETED | CB_MOLTEN | CB_ADDED); }
            hdr[i] = make_uint4(off, (u32)sz, sig, bits);
        } else {
            const uint4 h = hdr[i];
            if (C_DELETED(h.w)) continue;
            sz = (int)h.y; off = h.x; sig = h.z;
            if (sz <= 8) {
#pragma unroll
                for (int q = 0; q < 8; q++) r[q] = (q < sz) ? pool[off + q] : 0xFFFFFFFFu;
            }
        }
        u32 rk[4] = {0, 0, 0, 0};
        u32 first = 0, last = 0;
        if (sz <= 8) {
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (q < sz) {
                    const u32 lit = r[q];
                    if (q == 0) first = lit;
                    last = lit;
                    rk[q >> 1] |= atomicAdd(&cntS[lit >> shift], 1u) << ((q & 1) * 16);
                }
        } else {
            const u32* l = pool + off;
            first = l[0]; last = l[sz - 1];
            for (int q = 0; q < sz; q++) atomicAdd(&cntL[l[q] >> shift], 1u);
            if (sz >= (1 << 14)) atomicOr(flags, 8u);   // the list sort's folded key needs size < 2^14 (otsort.cu)
        }
        rk8[i] = make_uint4(rk[0], rk[1], rk[2], rk[3]);
        key[i] = make_uint4((u32)sz, first, last, sig);
    }
    __syncthreads();
    u32* row = cntMat + (size_t)blockIdx.x * NBp;
    for (u32 b = threadIdx.x; b < NBp; b += OT_T) row[b] = b < NB ? cntS[b] + cntL[b] : 0u;
}

// Run starts without atomics: the start of tile t's run in bucket b is bstart[b] + sum of the counts of the tiles before t.
// Column scan of the count matrix in three small steps: per-segment column sums, one CTA that turns them into bucket starts
// and segment bases, per-segment running sums written to a second matrix (the counts themselves stay: the partition needs both).
#define OT_SEG 64
__global__ void __launch_bounds__(256) k_ot_colsum(const u32* __restrict__ cntMat, u32 tiles, u32 NB, u32 NBp, u32* __restrict__ segSum) {
    const u32 b = blockIdx.x * 256 + threadIdx.x;
    if (b >= NB) return;
    const u32 per = (tiles + OT_SEG - 1) / OT_SEG;
    const u32 r0 = blockIdx.y * per, r1 = min(tiles, r0 + per);
    u32 s = 0;
    for (u32 r = r0; r < r1; r++) s += cntMat[(size_t)r * NBp + b];
    segSum[(size_t)blockIdx.y * NBp + b] = s;
}
// bucket totals -> exclusive scan (NB <= 8192) -> segment starts; segSum becomes the start of every segment's first run; one CTA
__global__ void __launch_bounds__(1024) k_ot_bscan(u32* __restrict__ segSum, u32 NB, u32 NBp, u32* __restrict__ bstart, u32* total) {
    __shared__ u32 wt[32];
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < NB; base += 1024) {
        const u32 b = base + threadIdx.x;
        u32 v = 0;
        if (b < NB) for (u32 sgm = 0; sgm < OT_SEG; sgm++) v += segSum[(size_t)sgm * NBp + b];
        const u32 incl = warpIncl(v);
        if ((threadIdx.x & 31u) == 31u) wt[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) { const u32 t = wt[threadIdx.x]; const u32 ti = warpIncl(t); wt[threadIdx.x] = ti - t; }
        __syncthreads();
        const u32 excl = carry + wt[threadIdx.x >> 5] + incl - v;
        if (b < NB) {
            bstart[b] = excl;
            u32 run = excl;
            for (u32 sgm = 0; sgm < OT_SEG; sgm++) { const u32 t = segSum[(size_t)sgm * NBp + b]; segSum[(size_t)sgm * NBp + b] = run; run += t; }
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { bstart[NB] = carry; *total = carry; }
}
__global__ void __launch_bounds__(256) k_ot_colfix(const u32* __restrict__ cntMat, u32 tiles, u32 NB, u32 NBp, const u32* __restrict__ segBase,
                                                   u32* __restrict__ runMat) {
    const u32 b = blockIdx.x * 256 + threadIdx.x;
    if (b >= NB) return;
    const u32 per = (tiles + OT_SEG - 1) / OT_SEG;
    const u32 r0 = blockIdx.y * per, r1 = min(tiles, r0 + per);
    u32 run = segBase[(size_t)blockIdx.y * NBp + b];
    for (u32 r = r0; r < r1; r++) { runMat[(size_t)r * NBp + b] = run; run += cntMat[(size_t)r * NBp + b]; }
}

template <int CPT, int KEEP>
__global__ void __launch_bounds__(OT_T, 1) k_ot_part2(const uint4* __restrict__ hdr, const u32* __restrict__ pool, const uint4* __restrict__ rk8, u32 n,
                                                   u32 shift, u32 NB, u32 NBp, const u32* __restrict__ cntMat, const u32* __restrict__ runMat,
                                                   u32 stageCap, uint2* __restrict__ pairs) {
    extern __shared__ u32 sm[];
    u32* tileOff = sm;              // [NB + 1] start of the bucket inside the staged tile
    u32* delta = sm + NB + 1;       // [NB] run start in pairs[] minus tileOff
    u32* cntL2 = delta + NB;        // [NB] long-clause literals placed so far (tail of the run, downwards)
    uint2* stage = (uint2*)(sm + ((3 * NB + 2) & ~1u));
    __shared__ u32 warpTot[32];
    __shared__ u32 tileTotal, nonEmpty;
    const u32 tile0 = blockIdx.x * (OT_T * CPT);
    if (threadIdx.x == 0) nonEmpty = 0;
    u32 off[CPT], sz[CPT]; uint4 rk[CPT];
    bool longHere = false;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const u32 i = tile0 + k * OT_T + threadIdx.x;
        sz[k] = 0; off[k] = 0; rk[k] = make_uint4(0, 0, 0, 0);
        if (i < n) {
            const uint4 h = hdr[i];
            if (!C_DELETED(h.w)) { off[k] = h.x; sz[k] = h.y; rk[k] = rk8[i]; longHere |= h.y > 8u; }
        }
    }
    const int anyLong = __syncthreads_or(longHere);
    if (anyLong) for (u32 b = threadIdx.x; b < NB; b += OT_T) cntL2[b] = 0;
    // the first literals of every short clause: loads in flight while the runs are reserved
    u32 lk[CPT][KEEP];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const u32* l = pool + off[k];
#pragma unroll
        for (int q = 0; q < KEEP; q++) lk[k][q] = ((u32)q < sz[k] && sz[k] <= 8u) ? l[q] : 0u;
    }
    // this tile's rows: run lengths (counting pass) and run starts (column scan) - no reservation, no atomic
    const u32 per = (NB + OT_T - 1) / OT_T;   // consecutive buckets per thread, <= 8
    const u32 b0 = threadIdx.x * per;
    const u32* row = cntMat + (size_t)blockIdx.x * NBp;
    const u32* grow = runMat + (size_t)blockIdx.x * NBp;
    u32 cq[8], gq[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { const u32 b = b0 + q; const bool in = (u32)q < per && b < NB; cq[q] = in ? row[b] : 0u; gq[q] = in ? grow[b] : 0u; }
    u32 mine = 0, used = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { mine += cq[q]; used += cq[q] != 0u; }
    const u32 incl = warpIncl(mine);
    used = warpSum(used);
    if ((threadIdx.x & 31u) == 0 && used) atomicAdd(&nonEmpty, used);
    if ((threadIdx.x & 31u) == 31u) warpTot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const u32 t = warpTot[threadIdx.x];
        const u32 ti = warpIncl(t);
        warpTot[threadIdx.x] = ti - t;
        if (threadIdx.x == 31) { tileTotal = ti; tileOff[NB] = ti; }
    }
    __syncthreads();
    u32 run = warpTot[threadIdx.x >> 5] + incl - mine;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const u32 b = b0 + q;
        if ((u32)q < per && b < NB) { tileOff[b] = run; delta[b] = gq[q] - run; run += cq[q]; }
    }
    __syncthreads();
    // staging pays when the tile's runs are short (uniform random formulas); clause-local formulas already write long runs
    const bool staged = tileTotal <= stageCap && tileTotal < 12u * nonEmpty;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const u32 i = tile0 + k * OT_T + threadIdx.x;
        const u32* l = pool + off[k];
        if (sz[k] <= 8u) {
            const u32 rw[4] = {rk[k].x, rk[k].y, rk[k].z, rk[k].w};
#pragma unroll
            for (int q = 0; q < 8; q++)
                if ((u32)q < sz[k]) {
                    const u32 lit = q < KEEP ? lk[k][q < KEEP ? q : 0] : l[q];
                    const u32 b = lit >> shift;
                    const u32 pos = tileOff[b] + ((rw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu);
                    if (staged) stage[pos] = make_uint2(lit, i);
                    else pairs[pos + delta[b]] = make_uint2(lit, i);
                }
        } else
            for (u32 q = 0; q < sz[k]; q++) {
                const u32 lit = l[q];
                const u32 b = lit >> shift;
                const u32 pos = tileOff[b + 1] - 1u - atomicAdd(&cntL2[b], 1u);
                if (staged) stage[pos] = make_uint2(lit, i);
                else pairs[pos + delta[b]] = make_uint2(lit, i);
            }
    }
    if (!staged) return;
    __syncthreads();
    const u32 total = tileTotal;
    for (u32 t = threadIdx.x; t < total; t += OT_T) {   // (four-way unrolling measured no faster: profiles/r02_ab_c10.jsonl)
        const uint2 pr = stage[t];
        pairs[t + delta[pr.x >> shift]] = pr;
    }
}

// Version 3 of the partition: the same tile algorithm on PERSISTENT CTAs (one per SM) with the tiles software-pipelined.
// k_ot_part2 holds ~220 KB of shared memory, so one CTA lives on an SM and its phases run one after the other: header
// load -> literal load (dependent) -> row scan -> staging -> copy-out, and the next CTA cannot start before the last
// store has drained (ncu: long-scoreboard + barrier + drain = half of the stall cycles, LSU pipe 48 % busy).  Here the
// loads of tile i+1 are in flight while tile i is copied out: its two matrix rows come in through cp.async
// (global -> shared memory, no registers), headers and ranks are issued before the first half of the copy-out and the
// literals - which need the headers - before the second half; the registers of tile i are dead by then.
__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gsrc) {
    const u32 d = (u32)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int CPT, int KEEP>
__global__ void __launch_bounds__(OT_T, 1) k_ot_part3(const uint4* __restrict__ hdr, const u32* __restrict__ pool, const uint4* __restrict__ rk8, u32 n,
                                                   u32 shift, u32 NB, u32 NBp, const u32* __restrict__ cntMat, const u32* __restrict__ runMat,
                                                   u32 stageCap, u32 tiles, uint2* __restrict__ pairs) {
    extern __shared__ __align__(16) u32 sm3[];
    u32* rowC = sm3;                 // [NBp] this tile's run lengths   (cp.async target, 16-byte aligned)
    u32* rowG = sm3 + NBp;           // [NBp] this tile's run starts
    u32* tileOff = sm3 + 2 * NBp;    // [NB + 1]
    u32* delta = tileOff + NB + 1;  // [NB]
    u32* cntL2 = delta + NB;        // [NB]
    uint2* stage = (uint2*)(sm3 + ((2 * NBp + 3 * NB + 2) & ~1u));
    __shared__ u32 warpTot[32];
    __shared__ u32 tileTotal, nonEmpty;
    const u32 per = (NB + OT_T - 1) / OT_T;   // consecutive buckets per thread, <= 8
    const u32 b0 = threadIdx.x * per;
    u32 tile = blockIdx.x;
    if (tile >= tiles) return;
    u32 off[CPT], sz[CPT]; uint4 rk[CPT]; u32 lk[CPT][KEEP];
#define PART3_LOAD_HDR(T_)                                                                                        \
    _Pragma("unroll") for (int k = 0; k < CPT; k++) {                                                             \
        const u32 i = (T_) * (OT_T * CPT) + k * OT_T + threadIdx.x;                                               \
        sz[k] = 0; off[k] = 0; rk[k] = make_uint4(0, 0, 0, 0);                                                    \
        if (i < n) { const uint4 h = hdr[i]; if (!C_DELETED(h.w)) { off[k] = h.x; sz[k] = h.y; rk[k] = rk8[i]; } } \
    }
#define PART3_LOAD_ROWS(T_)                                                                                       \
    do {                                                                                                          \
        const u32* rc = cntMat + (size_t)(T_) * NBp; const u32* rg = runMat + (size_t)(T_) * NBp;                 \
        for (u32 w = threadIdx.x * 4; w < NBp; w += OT_T * 4) { cpAsync16(rowC + w, rc + w); cpAsync16(rowG + w, rg + w); } \
        cpAsyncCommit();                                                                                          \
    } while (0)
#define PART3_LOAD_LITS()                                                                                         \
    _Pragma("unroll") for (int k = 0; k < CPT; k++) {                                                             \
        const u32* l = pool + off[k];                                                                             \
        _Pragma("unroll") for (int q = 0; q < KEEP; q++) lk[k][q] = ((u32)q < sz[k] && sz[k] <= 8u) ? l[q] : 0u;  \
    }
    PART3_LOAD_HDR(tile);
    PART3_LOAD_ROWS(tile);
    PART3_LOAD_LITS();
    for (;;) {
        if (threadIdx.x == 0) nonEmpty = 0;
        bool longHere = false;
#pragma unroll
        for (int k = 0; k < CPT; k++) longHere |= sz[k] > 8u;
        cpAsyncWaitAll();
        const int anyLong = __syncthreads_or(longHere);   // rows of this tile visible to every thread
        if (anyLong) for (u32 b = threadIdx.x; b < NB; b += OT_T) cntL2[b] = 0;
        u32 cq[8];
        u32 mine = 0, used = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { const u32 b = b0 + q; cq[q] = ((u32)q < per && b < NB) ? rowC[b] : 0u; mine += cq[q]; used += cq[q] != 0u; }
        const u32 incl = warpIncl(mine);
        used = warpSum(used);
        if ((threadIdx.x & 31u) == 0 && used) atomicAdd(&nonEmpty, used);
        if ((threadIdx.x & 31u) == 31u) warpTot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const u32 t = warpTot[threadIdx.x];
            const u32 ti = warpIncl(t);
            warpTot[threadIdx.x] = ti - t;
            if (threadIdx.x == 31) { tileTotal = ti; tileOff[NB] = ti; }
        }
        __syncthreads();
        u32 run = warpTot[threadIdx.x >> 5] + incl - mine;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const u32 b = b0 + q;
            if ((u32)q < per && b < NB) { tileOff[b] = run; delta[b] = rowG[b] - run; run += cq[q]; }
        }
        __syncthreads();
        const u32 total = tileTotal;
        const bool staged = total <= stageCap && total < 12u * nonEmpty;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const u32 i = tile * (OT_T * CPT) + k * OT_T + threadIdx.x;
            const u32* l = pool + off[k];
            if (sz[k] <= 8u) {
                const u32 rw[4] = {rk[k].x, rk[k].y, rk[k].z, rk[k].w};
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if ((u32)q < sz[k]) {
                        const u32 lit = q < KEEP ? lk[k][q < KEEP ? q : 0] : l[q];
                        const u32 b = lit >> shift;
                        const u32 pos = tileOff[b] + ((rw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu);
                        if (staged) stage[pos] = make_uint2(lit, i);
                        else pairs[pos + delta[b]] = make_uint2(lit, i);
                    }
            } else
                for (u32 q = 0; q < sz[k]; q++) {
                    const u32 lit = l[q];
                    const u32 b = lit >> shift;
                    const u32 pos = tileOff[b + 1] - 1u - atomicAdd(&cntL2[b], 1u);
                    if (staged) stage[pos] = make_uint2(lit, i);
                    else pairs[pos + delta[b]] = make_uint2(lit, i);
                }
        }
        __syncthreads();   // the stage is complete; rowC / rowG and this tile's registers are free
        const u32 next = tile + gridDim.x;
        const bool more = next < tiles;
        if (more) { PART3_LOAD_ROWS(next); PART3_LOAD_HDR(next); }
        asm volatile("" ::: "memory");
        const u32 half = staged ? ((total >> 1) & ~(u32)(OT_T - 1)) : 0u;
        u32 t = threadIdx.x;
        for (; t < half; t += OT_T) { const uint2 pr = stage[t]; pairs[t + delta[pr.x >> shift]] = pr; }
        asm volatile("" ::: "memory");
        if (more) { PART3_LOAD_LITS(); }
        asm volatile("" ::: "memory");
        if (staged) for (; t < total; t += OT_T) { const uint2 pr = stage[t]; pairs[t + delta[pr.x >> shift]] = pr; }
        if (!more) break;
        tile = next;
        __syncthreads();   // every reader of stage / delta is done before the next tile rewrites them
    }
#undef PART3_LOAD_HDR
#undef PART3_LOAD_ROWS
#undef PART3_LOAD_LITS
}

// per-literal histogram of one bucket = hist[] (and, scanned, otStart[]) of its literal range: counted from the bucket's
// pairs with shared-memory atomics - the global reductions of the version-1 histogram pass are gone.  Light (W counters),
// several CTAs per SM; oversized buckets just loop longer.  k_ot_place / k_ot_place_big then run unchanged.
#define LITHIST_T 512
__global__ void __launch_bounds__(LITHIST_T) k_ot_lithist(const uint2* __restrict__ pairs, const u32* __restrict__ bstart, u32 ND, u32 shift,
                                                          u32* __restrict__ hist, u32* __restrict__ otStart) {
    extern __shared__ u32 cnt[];   // [W]
    __shared__ u32 wt[32];
    __shared__ u32 carry;
    const u32 W = 1u << shift;
    const u32 lit0 = blockIdx.x << shift;
    const u32 nl = min(lit0 + W, ND) - lit0;
    const u32 p0 = bstart[blockIdx.x], p1 = bstart[blockIdx.x + 1];
    for (u32 k = threadIdx.x; k < W; k += LITHIST_T) cnt[k] = 0;
    if (threadIdx.x == 0) carry = p0;
    __syncthreads();
    u32 j = p0 + threadIdx.x;
    for (; j + 3 * LITHIST_T < p1; j += 4 * LITHIST_T) {
        u32 l4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) l4[k] = pairs[j + k * LITHIST_T].x;
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(&cnt[l4[k] - lit0], 1u);
    }
    for (; j < p1; j += LITHIST_T) atomicAdd(&cnt[pairs[j].x - lit0], 1u);
    __syncthreads();
    for (u32 base = 0; base < nl; base += LITHIST_T) {
        const u32 k = base + threadIdx.x;
        const u32 v = k < nl ? cnt[k] : 0u;
        const u32 incl = warpIncl(v);
        if ((threadIdx.x & 31u) == 31u) wt[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) { const u32 t = threadIdx.x < LITHIST_T / 32 ? wt[threadIdx.x] : 0u; const u32 ti = warpIncl(t); if (threadIdx.x < LITHIST_T / 32) wt[threadIdx.x] = ti - t; }
        __syncthreads();
        const u32 excl = carry + wt[threadIdx.x >> 5] + incl - v;
        if (k < nl) { hist[lit0 + k] = v; otStart[lit0 + k] = excl; }
        __syncthreads();
        if (threadIdx.x == LITHIST_T - 1) carry = excl + v;
        __syncthreads();
    }
}

// bucket = 2^shift consecutive literals, sized so that an average bucket fills at most ~70 % of a placement window; at most
// 8192 buckets.  Fixed BEFORE the counting pass: pass A, the partition and the placement must agree on it.
static void otChooseShape(Ctx* c, u64 numLiterals, u64 numClauses, u32 nSlots) {
    u32 shift = 6;
    while (shift < 12 && ((u64)numLiterals << (shift + 1)) / c->ND <= PLACE_WINDOW * 7 / 10) shift++;
    while (shift < 15 && ((c->ND + (1u << shift) - 1) >> shift) > 8192) shift++;
    c->otShift = shift; c->otNB = (c->ND + (1u << shift) - 1) >> shift;
    c->otNBp = (c->otNB + 3u) & ~3u;
    c->otCPT = numLiterals <= 3 * numClauses ? 5 : 3;   // short clauses: more of them per tile
    c->otTiles = divup(nSlots, OT_T * c->otCPT);
}
bool otBuildV2();
static bool otV2() { return otBuildV2(); }
bool otBuildV2() { static const int v = getenv("SIGMA_OT_V2") ? atoi(getenv("SIGMA_OT_V2")) : 1; return v != 0; }

static void launchCountPass(Ctx* c, bool awaken, u32 n, u64 numLiterals, u64 numClauses) {
    otChooseShape(c, numLiterals, numClauses, n);
    if (!c->attrOT2) {
        cudaFuncSetAttribute(k_ot_count<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_count<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_count<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_count<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_part2<3, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_part2<5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_part3<3, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_part3<5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_place, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (PLACE_WINDOW + (1 << 12)));
        cudaFuncSetAttribute(k_ot_lithist, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << 15);
        c->attrOT2 = true;
    }
    const size_t smem = 8 * (size_t)c->otNB;
    const u32 NB = c->otNB, NBp = c->otNBp, sh = c->otShift;
#define OT_COUNT_ARGS c->inLits, c->inOffs, c->inMeta, c->L0, c->hdr[c->cur], c->pool[c->cur], n, c->ND, sh, NB, NBp, c->rk8, c->cntMat, c->key, &c->dc->flags
    if (awaken) {
        if (c->otCPT == 5) LAUNCH(c, (k_ot_count<5, true>), c->otTiles, OT_T, smem, OT_COUNT_ARGS);
        else LAUNCH(c, (k_ot_count<3, true>), c->otTiles, OT_T, smem, OT_COUNT_ARGS);
        KB(c, 8.0 * n + 4.0 * numLiterals + (c->inMeta ? 4.0 * n : 0.0) + 16.0 * n + 4.0 * numLiterals + 32.0 * n + 4.0 * (double)c->otTiles * NBp);
    } else {
        if (c->otCPT == 5) LAUNCH(c, (k_ot_count<5, false>), c->otTiles, OT_T, smem, OT_COUNT_ARGS);
        else LAUNCH(c, (k_ot_count<3, false>), c->otTiles, OT_T, smem, OT_COUNT_ARGS);
        KB(c, 16.0 * n + 4.0 * numLiterals + 32.0 * numClauses + 4.0 * (double)c->otTiles * NBp);   // headers + literals in, keys + ranks + count rows out
    }
#undef OT_COUNT_ARGS
}

static void launchScatter2(Ctx* c, u32 n) {
    const u32 NB = c->otNB, NBp = c->otNBp, shift = c->otShift, tiles = c->otTiles;
    u32* segSum = c->otSeg;
    LAUNCH(c, k_ot_colsum, dim3(divup(NB, 256), OT_SEG), 256, 0, c->cntMat, tiles, NB, NBp, segSum);
    KB(c, 4.0 * (double)tiles * NBp);
    LAUNCH(c, k_ot_bscan, 1, 1024, 0, segSum, NB, NBp, c->bstart, c->otStart + c->ND);
    LAUNCH(c, k_ot_colfix, dim3(divup(NB, 256), OT_SEG), 256, 0, c->cntMat, tiles, NB, NBp, segSum, c->runMat);
    KB(c, 8.0 * (double)tiles * NBp);
    // shared memory of k_ot_part2: 3 words per bucket + the stage (whatever is left of ~220 KB, at most PART_STAGE pairs)
    const size_t partFixed = 4 * (((size_t)3 * NB + 2) & ~(size_t)1) + 16;
    const u32 stageCap = partFixed + 8 * (size_t)PART_STAGE <= 220 * 1024 ? PART_STAGE : (u32)((220 * 1024 - partFixed) / 8);
    // version 3 (persistent, software-pipelined tiles) needs two more rows of shared memory; it is used when its stage still
    // holds a whole tile of the loaded formula, otherwise version 2 (SIGMA_OT_PART3=0 forces version 2 for A/B runs)
    static const int part3 = getenv("SIGMA_OT_PART3") ? atoi(getenv("SIGMA_OT_PART3")) : 1;
    const size_t p3Fixed = 4 * (((size_t)2 * NBp + 3 * NB + 2) & ~(size_t)1) + 16;
    const u32 p3Cap = p3Fixed + 8 * (size_t)PART_STAGE <= 220 * 1024 ? PART_STAGE : (p3Fixed < 220 * 1024 ? (u32)((220 * 1024 - p3Fixed) / 8) : 0u);
    const u64 tileLits = c->numClauses ? (u64)OT_T * c->otCPT * c->numLiterals / c->numClauses : 0;
    if (part3 && tiles > 148 && (p3Cap >= stageCap || tileLits + tileLits / 8 <= p3Cap)) {
        if (c->otCPT == 5)
            LAUNCH(c, (k_ot_part3<5, 3>), 148, OT_T, p3Fixed + 8 * (size_t)p3Cap, c->hdr[c->cur], c->pool[c->cur], c->rk8, n, shift, NB, NBp, c->cntMat, c->runMat,
                   p3Cap, tiles, c->otPairs);
        else
            LAUNCH(c, (k_ot_part3<3, 5>), 148, OT_T, p3Fixed + 8 * (size_t)p3Cap, c->hdr[c->cur], c->pool[c->cur], c->rk8, n, shift, NB, NBp, c->cntMat, c->runMat,
                   p3Cap, tiles, c->otPairs);
    } else if (c->otCPT == 5)
        LAUNCH(c, (k_ot_part2<5, 3>), tiles, OT_T, partFixed + 8 * (size_t)stageCap, c->hdr[c->cur], c->pool[c->cur], c->rk8, n, shift, NB, NBp, c->cntMat, c->runMat,
               stageCap, c->otPairs);
    else
        LAUNCH(c, (k_ot_part2<3, 5>), tiles, OT_T, partFixed + 8 * (size_t)stageCap, c->hdr[c->cur], c->pool[c->cur], c->rk8, n, shift, NB, NBp, c->cntMat, c->runMat,
               stageCap, c->otPairs);
    KB(c, 32.0 * n + 4.0 * c->numLiterals + 8.0 * (double)tiles * NBp + 8.0 * c->numLiterals);   // headers + ranks + literals + the tile's two rows in, pairs out
    LAUNCH(c, k_ot_lithist, NB, LITHIST_T, (size_t)4 << shift, c->otPairs, c->bstart, c->ND, shift, c->hist, c->otStart);
    KB(c, 8.0 * c->numLiterals + 8.0 * c->ND);   // pairs in, hist + list starts out
    // placement: the version-1 kernels (list cursors + the bucket's occurs[] window in shared memory; work units for oversized buckets)
    u32 window = shift <= 12 ? PLACE_WINDOW : 0;
    if (const char* w = getenv("SIGMA_OT_WINDOW")) { const u32 v = (u32)atoi(w); if (v < window) window = v; }   // tests: force the work-unit path
    const size_t placeSmem = shift <= 12 ? 4 * ((size_t)PLACE_WINDOW + (1u << shift)) : (size_t)4 << shift;
    u32* nBig = &c->dc->scratch[7];
    cudaMemsetAsync(nBig, 0, 4, c->stream);
    // measured (profiles/r02_ab_tma_c11.jsonl, identical results): 0.373 vs 0.324 ms on cfg2, 1.45 vs 1.29 ms on cfg3 - the plain kernel
    // already keeps 4 x 1024 pair loads in flight per SM, the ring adds a CTA barrier per 16 KB chunk.  Opt-in.
    static const int placeTma = getenv("SIGMA_OT_TMA") ? atoi(getenv("SIGMA_OT_TMA")) : 0;
    if (placeTma && shift <= 12) {
        if (!c->attrTma) {
            cudaFuncSetAttribute(k_ot_place_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (PLACE_WINDOW + 8 + (1 << 12)) + 16 * PLACE_CH + 64);
            c->attrTma = true;
        }
        const size_t tmaSmem = 4 * ((size_t)PLACE_WINDOW + 8 + (1u << shift)) + 16 * (size_t)PLACE_CH + 64;
        LAUNCH(c, k_ot_place_tma, NB, PLACE_THREADS, tmaSmem, c->otPairs, c->otStart, c->ND, shift, window, c->otSize, c->occurs, c->otBig, nBig);
    } else
        LAUNCH(c, k_ot_place, NB, PLACE_THREADS, placeSmem, c->otPairs, c->otStart, c->ND, shift, window, c->otSize, c->occurs, c->otBig, nBig);
    KB(c, 8.0 * c->numLiterals + 4.0 * c->numLiterals + 8.0 * c->ND);   // pairs in, list entries out, list bounds
    LAUNCH(c, k_ot_place_big, 148 * 2, PLACE_THREADS, 0, c->otPairs, c->otStart, c->ND, shift, c->otSize, c->occurs, c->otBig, nBig);
}

// ------------------------------------------------------------------ live counts
// algorithmic bytes: read 16C (headers only)
__global__ void k_count_reset(DevCounters* dc) { dc->liveCls = 0; dc->liveLits = 0; }

__global__ void k_count(const uint4* __restrict__ hdr, DevCounters* dc) {
    const u32 n = dc->numCls;
    u32 nc = 0, nl = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (!C_DELETED(h.w)) nc++, nl += h.y;
    }
    nc = warpSum(nc); nl = warpSum(nl);
    __shared__ u32 sc[32], sl[32];
    const u32 w = threadIdx.x >> 5, l = threadIdx.x & 31u;
    if (l == 0) sc[w] = nc, sl[w] = nl;
    __syncthreads();
    if (w == 0) {
        nc = (l < (blockDim.x >> 5)) ? sc[l] : 0;
        nl = (l < (blockDim.x >> 5)) ? sl[l] : 0;
        nc = warpSum(nc); nl = warpSum(nl);
        if (l == 0 && (nc | nl)) { atomicAdd(&dc->liveCls, nc); atomicAdd((unsigned long long*)&dc->liveLits, (unsigned long long)nl); }
    }
}

void launchCount(Ctx* c) {
    LAUNCH(c, k_count_reset, 1, 1, 0, c->dc);
    LAUNCH(c, k_count, 148 * 4, 256, 0, c->hdr[c->cur], c->dc);
    KB(c, 16.0 * c->hdc->numCls);
}

// ------------------------------------------------------------------ GC compaction
// algorithmic bytes: read 16C_all + 4L_live, write 16C' + 4L' (+ two scans over C_all)
__global__ void k_gc_flags(const uint4* __restrict__ hdr, u32 n, u32* __restrict__ fCls, u32* __restrict__ fLits) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        const bool live = !C_DELETED(h.w);
        fCls[i] = live ? 1u : 0u;
        fLits[i] = live ? h.y : 0u;
    }
}

__global__ void k_gc_copy(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n,
                          const u32* __restrict__ pCls, const u32* __restrict__ pLits,
                          uint4* __restrict__ nhdr, u32* __restrict__ npool) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        const u32 j = pCls[i], off = pLits[i];
        const u32* s = pool + h.x;
        u32* d = npool + off;
        for (u32 k = 0; k < h.y; k++) d[k] = s[k];
        nhdr[j] = make_uint4(off, h.y, h.z, h.w);
    }
}

__global__ void k_gc_finish(DevCounters* dc, const u32* totCls, const u32* totLits) {
    dc->numCls = *totCls;
    dc->poolUsed = *totLits;
    dc->dataSize = (u64)NBUCKETS * (*totCls) + (*totLits);
}

// cuMM::compactCNF (recycle.cu:60-105): order preserving, storage shrunk to the current sizes
void launchGC(Ctx* c) {
    const u32 n = c->hdc->numCls;
    if (!n) return;
    const int src = c->cur, dst = 1 - c->cur;
    LAUNCH(c, k_gc_flags, gridFor(n, 256), 256, 0, c->hdr[src], n, c->flagA, c->flagB);
    KB(c, 24.0 * n);
    u32* tot = c->dc->scratch;
    scanExclusiveU32(c, c->flagA, c->flagA, n, 0, tot);
    scanExclusiveU32(c, c->flagB, c->flagB, n, 0, tot + 1);
    LAUNCH(c, k_gc_copy, gridFor(n, 256), 256, 0, c->hdr[src], c->pool[src], n, c->flagA, c->flagB, c->hdr[dst], c->pool[dst]);
    KB(c, 24.0 * n + 16.0 * c->numClauses + 8.0 * c->numLiterals);   // headers + scan values in, live literals in and out, live headers out
    LAUNCH(c, k_gc_finish, 1, 1, 0, c->dc, tot, tot + 1);
    c->cur = dst;
}

// ------------------------------------------------------------------ store
__global__ void k_store_arrays(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n,
                               const u32* __restrict__ pCls, const u32* __restrict__ pLits,
                               u32* __restrict__ oBits, u32* __restrict__ oSig, u64* __restrict__ oOffs, u32* __restrict__ oLits,
                               const u32* totCls, const u32* totLits) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        const u32 j = pCls[i], off = pLits[i];
        const u32* s = pool + h.x;
        for (u32 k = 0; k < h.y; k++) oLits[off + k] = s[k];
        oBits[j] = h.w;
        if (oOffs) { oSig[j] = h.z; oOffs[j] = off; }
        else oSig[j] = h.y;   // compact form: the size takes the place of the signature, no offsets
    }
    if (oOffs && blockIdx.x == 0 && threadIdx.x == 0) oOffs[*totCls] = *totLits;
}

// the reference's record stream: {bits, sig, size, lits...} at refs[j] (cnf.cuh:82-97)
__global__ void k_store_sclause(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n,
                                const u32* __restrict__ pCls, const u32* __restrict__ pLits,
                                u32* __restrict__ data, u64* __restrict__ refs) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        const u32 j = pCls[i];
        const u64 r = (u64)pLits[i] + (u64)NBUCKETS * j;
        u32* d = data + r;
        d[0] = h.w; d[1] = h.z; d[2] = h.y;
        const u32* s = pool + h.x;
        for (u32 k = 0; k < h.y; k++) d[3 + k] = s[k];
        refs[j] = r;
    }
}

// -aggresivesort (cacheCNF, cnf.cu:232-233): thrust::stable_sort of the refs with OLIST_CMP, i.e. the
// clauses leave ordered by (size, first literal, last literal, signature, ref).  Done as an LSD chain
// of stable radix sorts over the four key words (the initial order is the ref order), then the output
// position and literal offset of every clause are scattered back so that the store kernels run as usual.
__global__ void k_as_init(const uint4* __restrict__ hdr, u32 n, const u32* __restrict__ pCls, u32* __restrict__ order) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!C_DELETED(hdr[i].w)) order[pCls[i]] = i;
}
__global__ void k_as_word(const uint4* __restrict__ hdr, const u32* __restrict__ pool, const u32* __restrict__ order, u32 m, int word,
                          u32* __restrict__ keys) {
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const uint4 h = hdr[order[j]];
        keys[j] = word == 3 ? h.z : word == 2 ? (h.y ? pool[h.x + h.y - 1] : 0u) : word == 1 ? (h.y ? pool[h.x] : 0u) : h.y;
    }
}
__global__ void k_as_sizes(const uint4* __restrict__ hdr, const u32* __restrict__ order, u32 m, u32* __restrict__ sz) {
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) sz[j] = hdr[order[j]].y;
}
__global__ void k_as_back(const u32* __restrict__ order, const u32* __restrict__ offS, u32 m, u32* __restrict__ pCls, u32* __restrict__ pLits) {
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) { const u32 i = order[j]; pCls[i] = j; pLits[i] = offS[j]; }
}
static int aggressiveOrder(Ctx* c, u32 n) {
    // live count first (host): the radix launches are sized by it
    u32 m = 0;
    cudaError_t e = cudaMemcpyAsync(&m, c->dc->scratch, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return -(int)e;
    if (m < 2) return 0;
    u32* base = (u32*)c->otPairs;   // free outside the OT build: 8 capW bytes >= 5 capC words
    u32 *order = base, *order2 = base + c->capC, *keys = base + 2 * (size_t)c->capC, *keys2 = base + 3 * (size_t)c->capC, *offS = base + 4 * (size_t)c->capC;
    LAUNCH(c, k_as_init, gridFor(n, 256), 256, 0, c->hdr[c->cur], n, c->flagA, order);
    u32 litBits = 0;
    while (litBits < 32 && (c->ND >> litBits)) litBits++;
    const u32 bits[4] = {(c->hdc->flags & 8u) ? 32u : 14u, litBits, litBits, 32u};   // size, first, last, sig
    for (int word = 3; word >= 0; word--) {   // least significant key word first
        LAUNCH(c, k_as_word, gridFor(m, 256), 256, 0, c->hdr[c->cur], c->pool[c->cur], order, m, word, keys);
        radixSortPairs(c, keys, order, keys2, order2, m, bits[word]);
    }
    LAUNCH(c, k_as_sizes, gridFor(m, 256), 256, 0, c->hdr[c->cur], order, m, keys);
    scanExclusiveU32(c, keys, offS, m, 0, nullptr);
    LAUNCH(c, k_as_back, gridFor(m, 256), 256, 0, order, offS, m, c->flagA, c->flagB);
    return 0;
}

// Selects the live clauses into the staging buffers (the inactive CNF buffer); returns sizes.
int launchStore(Ctx* c, u64* nCls, u64* nLits, int form, bool writeBackOrder) {   // form: 0 arrays, 1 SCLAUSE records, 2 compact arrays
    const u32 n = c->hdc->numCls;
    const int src = c->cur, dst = 1 - c->cur;
    u32* tot = c->dc->scratch;
    *nCls = *nLits = 0;
    if (!n) return 0;
    LAUNCH(c, k_gc_flags, gridFor(n, 256), 256, 0, c->hdr[src], n, c->flagA, c->flagB);
    scanExclusiveU32(c, c->flagA, c->flagA, n, 0, tot);
    scanExclusiveU32(c, c->flagB, c->flagB, n, 0, tot + 1);
    if (writeBackOrder && c->o.aggr_cnf_sort && c->simpstate != SIGMA_CNFALLOC_FAIL && c->simpstate != SIGMA_OTALLOC_FAIL) {   // !reallocFailed()
        const int rc = aggressiveOrder(c, n);
        if (rc) return rc;
    }
    if (form == 3) {}   // selection only (flagA / flagB hold the output positions): sigma_continue copies into the input arrays
    else if (form == 1)
        LAUNCH(c, k_store_sclause, gridFor(n, 256), 256, 0, c->hdr[src], c->pool[src], n, c->flagA, c->flagB, c->pool[dst], c->flag64);
    else {
        u32* oBits = (u32*)c->hdr[dst];
        u32* oSig = oBits + c->capC;
        LAUNCH(c, k_store_arrays, gridFor(n, 256), 256, 0, c->hdr[src], c->pool[src], n, c->flagA, c->flagB, oBits, oSig,
               form == 2 ? (u64*)nullptr : c->flag64, c->pool[dst], tot, tot + 1);
    }
    u32 t[2];
    cudaError_t e = cudaMemcpyAsync(t, tot, 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { snprintf(c->err, sizeof c->err, "store: %s", cudaGetErrorString(e)); return -(int)e; }
    *nCls = t[0]; *nLits = t[1];
    return 0;
}
