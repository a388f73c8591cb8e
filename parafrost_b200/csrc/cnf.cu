// parafrost_b200/csrc/cnf.cu -- clause store kernels: awaken/prep fused with the counting pass of the occurrence-table
// build (bucket counts, ranks, sort keys), the table's radix partition + staged placement, live counts,
// garbage-collecting compaction, result store (+ the -aggresivesort write-back order).
//
// Reference behaviour being replaced (results must be identical):
//   prep_cnf_k            src/gpu/cnf.cu:45-53        sort literals, 32-bit signature        (k_ot_count<CPT, true>)
//   copy_if_k + histSimp  src/gpu/cnf.cu:33-43, histogram.cu:54-72  (thrust sort only to count!)  (k_ot_count, k_ot_lithist)
//   create_ot_k           src/gpu/occurrence.cu:50-62   (k_ot_part2 + k_ot_place / k_ot_place_big)
//   cnt_cls_lits          src/gpu/count.cu:83-106
//   scatter_k/compact_k   src/gpu/recycle.cu:34-105
//   cacheCNF              src/gpu/cnf.cu:200-237
//   thrust::stable_sort   src/gpu/cnf.cu:232-233      (-aggresivesort: aggressiveOrder)
// All of them stream the clause store once, coalesced; what bounds each is measured in DESIGN.md 7.
#include <cstdlib>

#include "common.cuh"

// ------------------------------------------------------------------ bucket shape of the occurrence-table build
// A bucket is a range of W consecutive literals, W = 2^sh or 3 * 2^sh: the 1.5x steps let an average bucket fill the
// placement window to 55-80 % instead of 35-70 % (uniform 5-SAT at 52 occurrences per literal: 768 literals per bucket
// instead of 512 - a third fewer buckets, a third longer runs in the partition).  `shape` = sh | three << 8.
__device__ __host__ __forceinline__ u32 bkW(u32 shape) { return ((shape >> 8) ? 3u : 1u) << (shape & 0xFFu); }
__device__ __forceinline__ u32 bkOf(u32 lit, u32 shape) { const u32 t = lit >> (shape & 0xFFu); return (shape >> 8) ? t / 3u : t; }
__device__ __forceinline__ u32 bkLit0(u32 b, u32 shape) { return ((shape >> 8) ? 3u * b : b) << (shape & 0xFFu); }

static void launchCountPass(Ctx* c, bool awaken, u32 n, u64 numLiterals, u64 numClauses);
static void launchScatter2(Ctx* c, u32 n);

// awaken + prep (prep_cnf_k, cnf.cu:45-53) fused with the first round's counting pass: k_ot_count<CPT, true>
void launchAwaken(Ctx* c) {
    if (!c->C0) return;
    launchCountPass(c, true, (u32)c->C0, c->L0, c->C0);
    c->histFresh = true;   // key[] and the count matrix describe the store until a kernel changes it (api.cu: buildOT)
}
// sort keys + bucket counts + ranks of the later rounds (copy_if_k + histSimp, histogram.cu:54-72): k_ot_count<CPT, false>
void launchHistKey(Ctx* c) {
    if (c->hdc->numCls) launchCountPass(c, false, c->hdc->numCls, c->numLiterals, c->numClauses);
}

// ------------------------------------------------------------------ occurrence lists: placement
// create_ot_k (occurrence.cu:50-62) appends every clause reference to the lists of its literals with one global atomic
// and one random 8-byte store per literal; random 4-byte stores over an occurs[] array far larger than L2 cost a DRAM
// sector each.  The lists are built MSD-radix style instead: (literal, clause) pairs partitioned by literal range
// (k_ot_count / k_ot_part2 below), then one CTA per bucket places the clause indices (k_ot_place).
#define PLACE_THREADS 1024
#define PLACE_SPLIT 64           // a bucket too large for the window is shared by up to 64 work units ...
#define PLACE_UNIT (32u << 10)   // ... of about this many pairs each
#define PLACE_WINDOW (50u << 10) // entries of a bucket's occurs[] window that can be staged in shared memory (200 KB)
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
    const u32 W = bkW(shift);
    u32* cur = smem;          // [W] list cursors
    u32* win = smem + W;      // [window] staged entries
    const u32 lit0 = bkLit0(blockIdx.x, shift);
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
    const u32 W = bkW(shift);
    const u32 nItems = *nBig;
    for (u32 item = blockIdx.x; item < nItems; item += gridDim.x) {
        const u32 code = big[item];
        const u32 b = code & 0x1FFFu, s = (code >> 13) & 0x3Fu, S = code >> 19;
        const u32 lit0 = bkLit0(b, shift);
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
    const u32 W = bkW(shift);
    u32* cur = smem;                                      // [W] list cursors
    u32* win = smem + W;                                  // [window + 8] staged entries, same 16-byte phase as occurs + p0
    uint2* ring = (uint2*)(smem + W + window + 8);        // [2][PLACE_CH]
    u64* bar = (u64*)(ring + 2 * PLACE_CH);               // [2]
    const u32 lit0 = bkLit0(blockIdx.x, shift);
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
        cudaMemsetAsync(c->hist, 0, (size_t)c->ND * 4, c->stream);
        return;
    }
    launchScatter2(c, n);
}

// ================================================================== occurrence-table build: count, scan, partition, histogram
// One atomic per literal twice, both in shared memory (round 1 paid three, one of them a global reduction: k_awaken at 0.33
// and k_ot_part at 0.25 of the copy roofline, profiles/r02_bench_cfg2_v2.json):
//   k_ot_count  (fused into awaken / the key pass, tile = 1024 x CPT clauses per CTA): one shared-memory atomic per literal on
//               the tile's BUCKET counter; the value it returns is the literal's rank inside (tile, bucket) and is kept -
//               8 x 16 bits per clause, one coalesced 16-byte store - together with the tile's row of bucket counts.
//               No global histogram, no global reductions.
//   k_ot_colsum / k_ot_bscan / k_ot_colfix: bucket totals = column sums of the count matrix, exclusive scan -> segment
//               starts, running column sums -> the start of every (tile, bucket) run.  No reservation atomics.
//   k_ot_part2  the same tile again: its two rows give run lengths and run starts, the stored ranks give every pair its
//               slot: NO atomics, no counter clearing; pairs are staged bucket by bucket and copied out with one table
//               look-up per pair.  Literals of clauses longer than 8 take the tail of the run through a second counter (rare).
//   k_ot_lithist one CTA per bucket: the per-literal histogram of the bucket's pairs (= hist[]; its local scan = otStart[])
//               with shared-memory atomics.
//   k_ot_place / k_ot_place_big (above): list cursors + the bucket's window in shared memory; work units for oversized buckets.
#define OT_T 1024

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
            if (C_DELETED(h.w)) { key[i] = make_uint4(0, 0, 0, 0); continue; }   // size 0 = no clause: k_ere_bloom streams the keys alone
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
                    rk[q >> 1] |= atomicAdd(&cntS[bkOf(lit, shift)], 1u) << ((q & 1) * 16);
                }
        } else {
            const u32* l = pool + off;
            first = l[0]; last = l[sz - 1];
            for (int q = 0; q < sz; q++) atomicAdd(&cntL[bkOf(l[q], shift)], 1u);
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
                    const u32 b = bkOf(lit, shift);
                    const u32 pos = tileOff[b] + ((rw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu);
                    if (staged) stage[pos] = make_uint2(lit, i);
                    else pairs[pos + delta[b]] = make_uint2(lit, i);
                }
        } else
            for (u32 q = 0; q < sz[k]; q++) {
                const u32 lit = l[q];
                const u32 b = bkOf(lit, shift);
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
        pairs[t + delta[bkOf(pr.x, shift)]] = pr;
    }
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
    const u32 W = bkW(shift);
    const u32 lit0 = bkLit0(blockIdx.x, shift);
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

// Shape of the build, fixed BEFORE the counting pass (counting pass, partition and placement must agree on it):
//   bucket width W = 2^sh or 3 * 2^sh literals - the largest with an average bucket at most 80 % of the placement window
//   (SIGMA_OT_FILL); at most 8192 buckets (bucket index < 2^13 in the work-unit codes, counters in shared memory).
//   Formulas with hot literal ranges (multiplier inputs: one bucket with 28x the mean) overflow the window whatever the
//   width, and a wider bucket only sends more pairs through the work-unit path: once a build of the loaded formula has
//   queued work units, the later ones go back to powers of two at <= 70 % of a 40 K-entry window (measured on cfg3 / cfg4,
//   profiles/r02_ab_c15.jsonl);
//   tile = 1024 x CPT clauses: 5 per thread for short clauses, else 3.
#define PART2_STAGE 26624u   // pairs of a tile staged in shared memory, at most (what 220 KB leave after the bucket tables decides)
static bool placeTmaOn() { static const int v = getenv("SIGMA_OT_TMA") ? atoi(getenv("SIGMA_OT_TMA")) : 0; return v != 0; }
// entries of a bucket's occurs[] window staged by the placement; the bulk-copy variant also holds a 32 KB ring
static u32 placeWindow() { return placeTmaOn() ? (40u << 10) : PLACE_WINDOW; }
#define PLACE_WMAX 6144u     // widest bucket whose list cursors fit beside the window
static size_t partFixedBytes(u32 NB) { return 4 * (((size_t)3 * NB + 2) & ~(size_t)1) + 16; }
// the stage holds the pairs of an average tile + 1/8 (a tile with more writes directly); no larger than needed: what the
// kernel does not take as shared memory stays L1 for its header / rank / literal loads (0.87 -> 0.80 ms on cfg2)
static u32 partStageCap(u32 NB, u64 tilePairs) {
    const size_t fixed = partFixedBytes(NB);
    const u64 room = fixed + 8 * (size_t)PART2_STAGE <= 220 * 1024 ? PART2_STAGE : (220 * 1024 - fixed) / 8;
    const u64 want = (tilePairs + tilePairs / 8 + 1023) & ~(u64)1023;
    return (u32)(want < room ? (want < 4096 ? 4096 : want) : room);
}
static void otChooseShape(Ctx* c, u64 numLiterals, u64 numClauses, u32 nSlots) {
    static const u32 fillPct = getenv("SIGMA_OT_FILL") ? (u32)atoi(getenv("SIGMA_OT_FILL")) : 80u;
    if (c->hdc->scratch[7]) c->otHot = true;   // the last build queued work units for oversized buckets
    const bool hot = c->otHot;
    const u64 limit = hot ? (u64)(40u << 10) * 7 / 10 : (u64)placeWindow() * fillPct / 100;
    u32 shape = 6;   // W = 64
    for (;;) {       // next wider candidate: 2^sh -> 3 * 2^(sh-1) -> 2^(sh+1)
        const u32 sh = shape & 0xFFu, three = shape >> 8;
        const u32 next = hot ? sh + 1 : (three ? (sh + 2) : ((sh - 1) | 0x100u));
        const u32 Wn = bkW(next);
        if (Wn > (hot ? 4096u : PLACE_WMAX) || numLiterals * Wn / c->ND > limit) break;
        shape = next;
    }
    while (bkW(shape) < (1u << 15) && (c->ND + bkW(shape) - 1) / bkW(shape) > 8192) shape = (shape >> 8) ? (shape & 0xFFu) + 2 : shape + 1;
    const u32 W = bkW(shape);
    c->otShift = shape; c->otNB = (c->ND + W - 1) / W;
    c->otNBp = (c->otNB + 3u) & ~3u;
    c->otCPT = numLiterals <= 3 * numClauses ? 5 : 3;   // short clauses: more of them per tile
    c->otTiles = divup(nSlots, OT_T * c->otCPT);
    c->otTilePairs = numClauses ? (u64)OT_T * c->otCPT * numLiterals / numClauses : 0;
}

static void launchCountPass(Ctx* c, bool awaken, u32 n, u64 numLiterals, u64 numClauses) {
    otChooseShape(c, numLiterals, numClauses, n);
    if (!c->attrOT2) {
        cudaFuncSetAttribute(k_ot_count<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_count<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_count<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_count<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        cudaFuncSetAttribute(k_ot_part2<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_part2<5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_ot_place, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (PLACE_WINDOW + PLACE_WMAX));
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
    const u32 W = bkW(shift);
    u32* segSum = c->otSeg;
    LAUNCH(c, k_ot_colsum, dim3(divup(NB, 256), OT_SEG), 256, 0, c->cntMat, tiles, NB, NBp, segSum);
    KB(c, 4.0 * (double)tiles * NBp);
    LAUNCH(c, k_ot_bscan, 1, 1024, 0, segSum, NB, NBp, c->bstart, c->otStart + c->ND);
    LAUNCH(c, k_ot_colfix, dim3(divup(NB, 256), OT_SEG), 256, 0, c->cntMat, tiles, NB, NBp, segSum, c->runMat);
    KB(c, 8.0 * (double)tiles * NBp);
    // shared memory of k_ot_part2: 3 words per bucket + the stage (whatever is left of ~220 KB, at most PART2_STAGE pairs)
    const size_t partFixed = partFixedBytes(NB);
    const u32 stageCap = partStageCap(NB, c->otTilePairs);
#define OT_PART_ARGS c->hdr[c->cur], c->pool[c->cur], c->rk8, n, shift, NB, NBp, c->cntMat, c->runMat, stageCap, c->otPairs
    // 3 literals per clause stay in registers between the sweeps, the others are re-read (L1): 5 of them cost spills under the
    // 64-register cap of a 1024-thread CTA (0.78 -> 0.73 ms on cfg2, profiles/r02_ab_c15.jsonl)
    if (c->otCPT == 5) LAUNCH(c, (k_ot_part2<5, 3>), tiles, OT_T, partFixed + 8 * (size_t)stageCap, OT_PART_ARGS);
    else LAUNCH(c, (k_ot_part2<3, 3>), tiles, OT_T, partFixed + 8 * (size_t)stageCap, OT_PART_ARGS);
#undef OT_PART_ARGS
    KB(c, 32.0 * n + 4.0 * c->numLiterals + 8.0 * (double)tiles * NBp + 8.0 * c->numLiterals);   // headers + ranks + literals + the tile's two rows in, pairs out
    LAUNCH(c, k_ot_lithist, NB, LITHIST_T, (size_t)4 * W, c->otPairs, c->bstart, c->ND, shift, c->hist, c->otStart);
    KB(c, 8.0 * c->numLiterals + 8.0 * c->ND);   // pairs in, hist + list starts out
    // placement: list cursors + the bucket's occurs[] window in shared memory; work units for oversized buckets
    // (hot formulas keep the 40 K window their buckets were sized for: shared memory the kernel does not take stays L1)
    const u32 winCap = c->otHot ? (40u << 10) : placeWindow();
    u32 window = W <= PLACE_WMAX ? winCap : 0;
    if (const char* w = getenv("SIGMA_OT_WINDOW")) { const u32 v = (u32)atoi(w); if (v < window) window = v; }   // tests: force the work-unit path
    const size_t placeSmem = W <= PLACE_WMAX ? 4 * ((size_t)winCap + W) : (size_t)4 * W;
    u32* nBig = &c->dc->scratch[7];
    cudaMemsetAsync(nBig, 0, 4, c->stream);
    // measured (profiles/r02_ab_tma_c11.jsonl, identical results): 0.373 vs 0.324 ms on cfg2, 1.45 vs 1.29 ms on cfg3 - the plain kernel
    // already keeps 4 x 1024 pair loads in flight per SM, the ring adds a CTA barrier per 16 KB chunk.  Opt-in.
    if (placeTmaOn() && W <= PLACE_WMAX) {
        if (!c->attrTma) {
            cudaFuncSetAttribute(k_ot_place_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * ((40 << 10) + 8 + PLACE_WMAX) + 16 * PLACE_CH + 64);
            c->attrTma = true;
        }
        const size_t tmaSmem = 4 * ((size_t)placeWindow() + 8 + W) + 16 * (size_t)PLACE_CH + 64;
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
    // four independent 16-byte loads per thread and iteration: a 0.09 ms kernel needs ~6 MB in flight to reach the copy bandwidth
    const u32 stride = gridDim.x * blockDim.x;
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; (u64)i + 3ull * stride < n; i += 4 * stride) {
        uint4 h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = hdr[i + k * stride];
#pragma unroll
        for (int k = 0; k < 4; k++) if (!C_DELETED(h[k].w)) nc++, nl += h[k].y;
    }
    for (; i < n; i += stride) {
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
    LAUNCH(c, k_count, 148 * 8, 256, 0, c->hdr[c->cur], c->dc);
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
