// parafrost_b200/csrc/lcve.cu -- the elected-variable schedule ("melting" an independent set).
//
// Replaces Solver::varReorder + Solver::LCVE (src/gpu/lcve.cu:280-398) in the reference's
// fixed-order mode (-no-lcvefast), where ONE GPU thread walks every candidate
// (lcve_k<<<1,1>>>, lcve.cu:64-101).  The serial walk elects, in ascending (score, var)
// order, every candidate that no earlier elected variable froze, and stops at the first
// unfrozen candidate that violates the occurrence bounds.  That is the lexicographically
// first maximal independent set of the prefix before the stop, so it is computed here in
// parallel and deterministically:
//   1. scores ps*ns (uint32 wrap) and candidate class            k_scores
//   2. stable LSD radix sort by score (ties keep ascending var)   k_radix_* (replaces thrust::sort + GPU_LCV_CMP, key.cuh:33-42)
//   3. rank-priority MIS rounds over a shrinking worklist, a candidate decides once no
//      lower-ranked undecided candidate shares a clause with it; "stoppers" (bound
//      violators) never block or freeze others, the first one that stays unfrozen cuts the
//      schedule.  The first round of a rank chunk is one streaming pass over the clauses
//      (k_mis_clauses / k_mis_first), later rounds probe a memorised blocker (k_mis_round);
//      a newly elected variable freezes its neighbours by push; a candidate with a clause
//      longer than lcveclausemax in its negative list only becomes MIS_HALF (common.cuh)
//   4. elected = decided-elected candidates below the cut, in rank order   k_elect_*
//   5. the first 12 frozen variables in the serial walk's order get their function-table
//      index (mapfrozen_k, lcve.cu:253-260; only indices < 12 are observable,
//      function.cuh:35,143)                                        k_frozen12
#include "common.cuh"

// ------------------------------------------------------------------ scores
// Per-variable election word vinfo[v] = rank << 5 | class << 3 | state : one 4-byte load per
// neighbour in the MIS rounds instead of three (state byte, class byte, rank word).
#define VI_STATE(w) ((w) & 7u)
#define VI_CLASS(w) (((w) >> 3) & 3u)
#define VI_RANK(w) ((w) >> 5)

// fast != 0 (-lcvefast, mis_init_k lcve.cu:150-173): the bound violators and the variables with an oversized clause
// (ovsFast, k_mark_oversize) are no candidates at all instead of cutting the walk
__global__ void k_scores(const u32* __restrict__ hist, const unsigned char* __restrict__ vstate,
                         const unsigned char* __restrict__ assumed, u32 V, u32 pmax, u32 nmax, u32 maxoccurs,
                         u32* __restrict__ keys, u32* __restrict__ vals, unsigned char* __restrict__ cstat, DevCounters* dc,
                         int fast, const unsigned char* __restrict__ ovsFast) {
    u32 kmax = 0;   // largest score: tells the host how many radix digits the sort needs
    for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < V; t += gridDim.x * blockDim.x) {
        const u32 v = t + 1;
        const u32 ps = hist[V2L(v)], ns = hist[V2L(v) | 1u];
        keys[t] = ps * ns;
        kmax = max(kmax, ps * ns);
        vals[t] = v;
        unsigned char cs = CS_NONE;
        if (!vstate[v] && !(assumed && assumed[v]) && (ps || ns)) {
            const bool stop = ps > maxoccurs || ns > maxoccurs || (ps >= pmax && ns >= nmax);
            if (!fast) cs = stop ? CS_STOP : CS_CAND;
            else cs = (stop || (ovsFast && ovsFast[v])) ? CS_NONE : CS_CAND;
        }
        cstat[v] = cs;
    }
    kmax = warpMax(kmax);
    if ((threadIdx.x & 31u) == 0 && kmax) atomicMax(&dc->scratch[6], kmax);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        dc->misStopRank = NOVAR; dc->numElected = 0; dc->wlNext = 0;
        dc->wlCnt[0] = dc->wlCnt[1] = dc->wlCnt[2] = 0; dc->firstStop = NOVAR;
    }
}

// ------------------------------------------------------------------ LSD radix sort (8-bit digits)
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const u32* __restrict__ keys, u32 n, u32 shift,
                                                            u32* __restrict__ ghist, u32 nblocks) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u32 base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const u32 i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    ghist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const u32* __restrict__ keysIn, const u32* __restrict__ valsIn,
                                                               u32* __restrict__ keysOut, u32* __restrict__ valsOut, u32 n,
                                                               u32 shift, const u32* __restrict__ ghist, u32 nblocks) {
    __shared__ u32 base[256];
    __shared__ u32 wc[RS_THREADS / 32][256];
    const u32 w = threadIdx.x >> 5, l = threadIdx.x & 31u;
    base[threadIdx.x] = ghist[threadIdx.x * nblocks + blockIdx.x];
    const u32 tile = blockIdx.x * RS_TILE;
    for (int k = 0; k < RS_ITEMS; k++) {
        for (u32 z = threadIdx.x; z < (RS_THREADS / 32) * 256; z += RS_THREADS) (&wc[0][0])[z] = 0;
        __syncthreads();
        const u32 i = tile + k * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        u32 key = 0, val = 0;
        if (valid) { key = keysIn[i]; val = valsIn[i]; }
        const u32 d = valid ? ((key >> shift) & 255u) : (256u + l);
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const u32 rankInWarp = __popc(peers & lanemaskLt());
        if (valid && rankInWarp == 0) wc[w][d] = __popc(peers);
        __syncthreads();
        {   // digit threadIdx.x: exclusive prefix over warps, advance the running base
            u32 run = base[threadIdx.x];
#pragma unroll
            for (int ww = 0; ww < RS_THREADS / 32; ww++) { const u32 t = wc[ww][threadIdx.x]; wc[ww][threadIdx.x] = run; run += t; }
            base[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const u32 pos = wc[w][d] + rankInWarp;
            keysOut[pos] = key; valsOut[pos] = val;
        }
        __syncthreads();
    }
}

// sorts (keys, vals) ascending by key, stable; only the digits below `bits` are sorted (all keys are
// < 2^bits) in an EVEN number of 8-bit passes, so that the result ends in (keys, vals)
void radixSortPairs(Ctx* c, u32* keys, u32* vals, u32* keys2, u32* vals2, u32 n, u32 bits) {
    const u32 nblocks = divup(n, RS_TILE);
    u32 *ki = keys, *vi = vals, *ko = keys2, *vo = vals2;
    u32 passes = (bits + 7) / 8;
    passes = (passes + 1) & ~1u;
    if (passes > 4) passes = 4;
    for (u32 shift = 0; shift < 8 * passes; shift += 8) {
        LAUNCH(c, k_radix_hist, nblocks, RS_THREADS, 0, ki, n, shift, c->radixHist, nblocks);
        KB(c, 4.0 * n);
        scanExclusiveU32(c, c->radixHist, c->radixHist, (u64)256 * nblocks, 0, nullptr);
        LAUNCH(c, k_radix_scatter, nblocks, RS_THREADS, 0, ki, vi, ko, vo, n, shift, c->radixHist, nblocks);
        KB(c, 16.0 * n);
        u32* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
}

__global__ void k_rank(const u32* __restrict__ eligible, const unsigned char* __restrict__ cstat, u32 V, u32* __restrict__ vinfo,
                       DevCounters* dc) {
    u32 firstStop = NOVAR;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < V; r += gridDim.x * blockDim.x) {
        const u32 v = eligible[r];
        const u32 cs = cstat[v];
        vinfo[v] = (r << 5) | (cs << 3) | (cs ? MIS_UNDECIDED : MIS_NONE);
        if (cs == CS_STOP && r < firstStop) firstStop = r;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) firstStop = min(firstStop, __shfl_xor_sync(0xffffffffu, firstStop, o));
    if ((threadIdx.x & 31u) == 0 && firstStop != NOVAR) atomicMin(&dc->firstStop, firstStop);
    if (blockIdx.x == 0 && threadIdx.x == 0) vinfo[0] = 0;
}

// ------------------------------------------------------------------ MIS
// worklist counters rotate over three slots: round k reads wlCnt[k%3], appends to wlCnt[(k+1)%3]
// and clears wlCnt[(k+2)%3], so several rounds can be queued without a host round-trip.
__global__ void k_mis_fill(const u32* __restrict__ eligible, const u32* __restrict__ vinfo, u32 rBegin, u32 rEnd,
                           u32* __restrict__ wl, DevCounters* dc, u32 slot) {
    for (u32 r = rBegin + blockIdx.x * blockDim.x + threadIdx.x; r < rEnd; r += gridDim.x * blockDim.x) {
        const u32 v = eligible[r];
        if (VI_STATE(vinfo[v]) == MIS_UNDECIDED) wl[warpAggInc(&dc->wlCnt[slot])] = v;   // candidates not yet frozen by an earlier chunk's push
    }
}

// one group of GS lanes per undecided candidate; lanes stride over the clauses of its two lists.
// blocker[v] remembers the lowest-ranked undecided candidate that shared a clause with v at its
// last full scan: while that one is undecided v stays blocked, when it is elected v is frozen -
// one 4-byte probe instead of walking the whole neighbourhood again; only when the blocker got
// frozen itself is the neighbourhood rescanned.  (Decisions can come a round later than with a
// full scan every round; the fixed point - the lexicographically first MIS - is the same.)
// the plain walk: right for the 8-lane groups of formulas with a handful of short clauses per
// variable, where batching only adds predicated loads (measured on cfg3: +45 %)
#define MIS_WALK_SIMPLE(GS_, V_, LANE_, NSIDES_, APPLY_)                                               \
    for (u32 side_ = 0; side_ < (NSIDES_); side_++) {                                                          \
        const u32 lit_ = V2L(V_) | side_;                                                              \
        const u32 n_ = otSize[lit_];                                                                   \
        const u32* list_ = occurs + otStart[lit_];                                                     \
        for (u32 j_ = (LANE_); j_ < n_; j_ += (GS_)) {                                                 \
            const uint4 h_ = hdr[list_[j_]];                                                           \
            if (C_DELETED(h_.w)) continue;                                                             \
            const u32 csize = h_.y; (void)csize;                                                       \
            pb_ += 20u + 8u * csize;   /* profile: list entry + header + literals + election words */  \
            const u32* l_ = pool + h_.x;                                                               \
            for (u32 k_ = 0; k_ < h_.y; k_++) { const u32 ul_ = l_[k_]; const u32 u = LABS(ul_); if (u != (V_)) { const bool upos = !LSIGN(ul_); (void)upos; const u32 wu = vinfo[u]; APPLY_ } } \
        }                                                                                              \
    }
// APPLY_ sees: u (neighbour variable), upos (its literal in the shared clause is positive), wu (its
// election word), csize (clause size), side_ (0: the clause is in v's positive list)
#define MIS_WALK_ANY(GS_, V_, LANE_, NSIDES_, APPLY_) \
    if constexpr ((GS_) == 32) { MIS_WALK(GS_, V_, LANE_, NSIDES_, APPLY_) } else { MIS_WALK_SIMPLE(GS_, V_, LANE_, NSIDES_, APPLY_) }

// An elected variable freezes its higher-ranked undecided neighbours at once (push) instead of
// letting each of them find out by rescanning its own neighbourhood: every concurrent writer of a
// neighbour's state word agrees on FROZEN (a neighbour of a just-elected variable can neither be
// elected nor become the live stopper in the same launch: it sees this variable undecided or elected).
// Walk over the neighbourhood of v (both occurrence lists, every literal of every live clause).  The
// chain list entry -> header -> literals -> election words is four dependent loads deep; with small
// worklists a MIS round is pure latency, so each lane keeps four clauses in flight (entries and
// headers loaded as a batch) and loads the literals and the election words of a clause as batches of 8.
// Both lists are walked as ONE index range (positive list first): the clauses of the two sides are in flight together
// instead of one side's chain after the other's - half the latency of a walk when the worklist is small.
#define MIS_WALK(GS_, V_, LANE_, NSIDES_, APPLY_)                                                      \
    {                                                                                                  \
        const u32 litP_ = V2L(V_);                                                                     \
        const u32 nP_ = otSize[litP_], nN_ = (NSIDES_) > 1u ? otSize[litP_ | 1u] : 0u;                 \
        const u32* listP_ = occurs + otStart[litP_];                                                   \
        const u32* listN_ = occurs + otStart[litP_ | 1u];                                              \
        const u32 n_ = nP_ + nN_;                                                                      \
        for (u32 j0_ = (LANE_); j0_ < n_; j0_ += 4 * (GS_)) {                                          \
            u32 ci_[4]; uint4 h_[4];                                                                   \
            _Pragma("unroll") for (int u_ = 0; u_ < 4; u_++) { const u32 j_ = j0_ + u_ * (GS_); ci_[u_] = j_ < n_ ? (j_ < nP_ ? listP_[j_] : listN_[j_ - nP_]) : NOVAR; } \
            _Pragma("unroll") for (int u_ = 0; u_ < 4; u_++) h_[u_] = ci_[u_] != NOVAR ? hdr[ci_[u_]] : make_uint4(0, 0, 0, CB_DELETED);  \
            _Pragma("unroll") for (int u_ = 0; u_ < 4; u_++) {                                         \
                if (C_DELETED(h_[u_].w)) continue;                                                     \
                const u32 side_ = (j0_ + u_ * (GS_)) >= nP_ ? 1u : 0u; (void)side_;                    \
                const u32 csize = h_[u_].y; (void)csize;                                               \
                pb_ += 20u + 8u * csize;                                                               \
                const u32* l_ = pool + h_[u_].x;                                                       \
                u32 lv_[8], wv_[8];                                                                    \
                _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) lv_[k_] = (u32)k_ < h_[u_].y ? l_[k_] : V2L(V_); /* literals */ \
                _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) wv_[k_] = LABS(lv_[k_]) != (V_) ? vinfo[LABS(lv_[k_])] : 0u;    \
                _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) if (LABS(lv_[k_]) != (V_)) { const u32 u = LABS(lv_[k_]), wu = wv_[k_]; const bool upos = !LSIGN(lv_[k_]); (void)upos; APPLY_ } \
                for (u32 k_ = 8; k_ < h_[u_].y; k_++) { const u32 ul_ = l_[k_]; const u32 u = LABS(ul_); if (u != (V_)) { const bool upos = !LSIGN(ul_); (void)upos; const u32 wu = vinfo[u]; APPLY_ } } \
            }                                                                                          \
        }                                                                                              \
    }

// An elected variable freezes its higher-ranked undecided neighbours at once (push) instead of
// letting each of them find out by rescanning its own neighbourhood: every concurrent writer of a
// neighbour's state word agrees on FROZEN (a neighbour of a just-elected variable can neither be
// elected nor become the live stopper in the same launch: it sees this variable undecided or elected).
template <int GS>
__device__ __forceinline__ u32 pushFreeze(u32 v, u32 r, u32 lane, const uint4* __restrict__ hdr, const u32* __restrict__ pool,
                                          const u32* __restrict__ otStart, const u32* __restrict__ otSize,
                                          const u32* __restrict__ occurs, u32* vinfo, u32 nsides) {
    // nsides = 1 for a MIS_HALF variable: only the clauses of its positive list freeze their variables
    u32 pb_ = 0;   // bytes this lane walked (kernel profile mode)
    MIS_WALK_ANY(GS, v, lane, nsides, { if (VI_STATE(wu) == MIS_UNDECIDED && VI_RANK(wu) > r) vinfo[u] = (wu & ~7u) | MIS_FROZEN; })
    return pb_;
}
// profile mode: the bytes walked by this thread -> *prof (one atomic per warp)
__device__ __forceinline__ void profAdd(unsigned long long* prof, unsigned long long pb) {
    if (!prof) return;
#pragma unroll
    for (int o = 16; o; o >>= 1) pb += __shfl_xor_sync(0xffffffffu, pb, o);
    if ((threadIdx.x & 31u) == 0 && pb) atomicAdd(prof, pb);
}

// push for the variables elected by k_mis_first (thread-per-variable there: no group to walk the lists)
template <int GS>
__global__ void __launch_bounds__(256) k_mis_push(const u32* __restrict__ list, const u32* __restrict__ count,
                                                  const uint4* __restrict__ hdr, const u32* __restrict__ pool,
                                                  const u32* __restrict__ otStart, const u32* __restrict__ otSize,
                                                  const u32* __restrict__ occurs, u32* vinfo, unsigned long long* prof) {
    const u32 n = *count;
    const u32 lane = threadIdx.x & (u32)(GS - 1);
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    unsigned long long pb = 0;
    for (u32 it = (blockIdx.x * blockDim.x + threadIdx.x) / GS; it < n; it += groupsPerGrid) {
        const u32 e = list[it];
        const u32 v = e & 0x7FFFFFFFu;   // bit 31: MIS_HALF
        pb += pushFreeze<GS>(v, VI_RANK(vinfo[v]), lane, hdr, pool, otStart, otSize, occurs, vinfo, (e >> 31) ? 1u : 2u);
    }
    profAdd(prof, pb);
}

template <int GS>
__global__ void __launch_bounds__(256) k_mis_round(u32* __restrict__ wl0, u32* __restrict__ wl1, u32 round, DevCounters* dc,
                                                   const uint4* __restrict__ hdr, const u32* __restrict__ pool,
                                                   const u32* __restrict__ otStart, const u32* __restrict__ otSize,
                                                   const u32* __restrict__ occurs, u32* vinfo, u32* __restrict__ blocker, int maxcsize,
                                                   unsigned long long* prof) {
    unsigned long long pb = 0;
    const u32* wlIn = (round & 1u) ? wl1 : wl0;
    u32* wlOut = (round & 1u) ? wl0 : wl1;
    const u32 nIn = dc->wlCnt[round % 3u];
    u32* outCnt = &dc->wlCnt[(round + 1u) % 3u];
    if (blockIdx.x == 0 && threadIdx.x == 0) dc->wlCnt[(round + 2u) % 3u] = 0;
    const u32 lane = threadIdx.x & (u32)(GS - 1);
    const u32 gbase = (threadIdx.x & 31u) & ~(u32)(GS - 1);
    const u32 gmask = GS == 32 ? 0xffffffffu : (((1u << (GS & 31)) - 1u) << gbase);
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    for (u32 it = (blockIdx.x * blockDim.x + threadIdx.x) / GS; it < nIn; it += groupsPerGrid) {
        const u32 v = wlIn[it];
        const u32 wv = vinfo[v];
        const u32 r = VI_RANK(wv);
        if (lane == 0) pb += 16;   // worklist entry, election word, blocker and its word
        if (VI_STATE(wv) != MIS_UNDECIDED) continue;   // frozen by an elected neighbour's push
        if (r > dc->misStopRank) continue;  // beyond the cut: never looked at by the serial walk
        const u32 b = blocker[v];
        if (b) {
            const u32 mb = VI_STATE(vinfo[b]);
            if (mb == MIS_UNDECIDED) { if (lane == 0) wlOut[atomicAdd(outCnt, 1u)] = v; continue; }
            if (mb == MIS_ELECTED) { if (lane == 0) vinfo[v] = (wv & ~7u) | MIS_FROZEN; continue; }
        }
        bool frozen = false, overP = false, overN = false;
        u32 bRank = NOVAR, bVar = 0;
        u32 pb_ = 0;
        MIS_WALK_ANY(GS, v, lane, 2u, {
            if ((int)csize > maxcsize) { if (side_ == 0) overP = true; else overN = true; }
            if (VI_RANK(wu) < r) {
                const u32 m = VI_STATE(wu);
                // an elected neighbour freezes v; a MIS_HALF one only through a clause of ITS positive list
                if (m == MIS_ELECTED || (m == MIS_HALF && upos)) frozen = true;
                else if (m == MIS_UNDECIDED && VI_CLASS(wu) == CS_CAND && VI_RANK(wu) < bRank) { bRank = VI_RANK(wu); bVar = u; }
            }
        })
        pb += pb_;
        frozen = __any_sync(gmask, frozen);
        overP = __any_sync(gmask, overP);
        overN = __any_sync(gmask, overN);
        u32 minRank = bRank;
#pragma unroll
        for (int o = GS / 2; o; o >>= 1) minRank = min(minRank, __shfl_xor_sync(gmask, minRank, o, GS));
        const bool blocked = minRank != NOVAR;
        if (blocked && !frozen && bRank == minRank) blocker[v] = bVar;   // ranks are unique: every writer holds the same variable
        if (lane == 0) {
            const u32 keep = wv & ~7u;
            if (frozen) vinfo[v] = keep | MIS_FROZEN;
            else if (!blocked) {
                if (VI_CLASS(wv) == CS_STOP) { vinfo[v] = keep | MIS_LIVESTOP; atomicMin(&dc->misStopRank, r); }
                // a clause longer than lcveclausemax makes depFreeze_d fail (lcve.cu:46-51): in the positive list -> not
                // elected, freezes nothing; only in the negative list -> not elected, the positive list's freezes stay
                else if (overP) vinfo[v] = keep | MIS_FROZEN;
                else if (overN) { vinfo[v] = keep | MIS_HALF; atomicAdd(&dc->scratch[8], 1u); }
                else vinfo[v] = keep | MIS_ELECTED;
            }
            else wlOut[atomicAdd(outCnt, 1u)] = v;
        }
        if (!frozen && !blocked && !overP && VI_CLASS(wv) == CS_CAND)   // just elected or MIS_HALF (uniform over the group)
            pb += pushFreeze<GS>(v, r, lane, hdr, pool, otStart, otSize, occurs, vinfo, overN ? 1u : 2u);
    }
    profAdd(prof, pb);
}

// ------------------------------------------------------------------ first MIS round, clause-centric
// The first round of a chunk is the expensive one: every candidate would walk its whole
// neighbourhood through the occurrence table (gathering headers and literals).  The same
// information falls out of ONE streaming pass over the clause store: per clause the lowest elected
// rank and the two lowest undecided candidate ranks; every undecided variable of the clause is
// frozen (an elected variable of lower rank shares the clause) or gets the lowest-ranked other
// undecided candidate min-reduced into nbr[].  k_mis_first then elects the local minima and
// queues the rest with their blocker for the vertex-centric rounds above.
// Algorithmic bytes: 16 C + 4 L (coalesced) + 4 L election words (L2 resident).
__global__ void __launch_bounds__(256) k_mis_clauses(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n, u32* vinfo,
                                                     u32* __restrict__ nbr, unsigned char* __restrict__ ovs, u32 H, int maxcsize) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        const u32* l = pool + h.x;
        u32 eMin = NOVAR, u1 = NOVAR, u2 = NOVAR;
        bool any = false;
        // literals and election words of the first 8 literals stay in registers for the second loop
        u32 lv[8], wv[8];
#pragma unroll
        for (u32 k = 0; k < 8; k++) if (k < h.y) lv[k] = l[k];
#pragma unroll
        for (u32 k = 0; k < 8; k++) if (k < h.y) wv[k] = vinfo[LABS(lv[k])];
        // a MIS_HALF variable freezes through the clauses of its positive list only (see MIS_HALF, common.cuh)
#define MC_SCAN(L_, W_) do { const u32 w = (W_); const u32 r = VI_RANK(w), st = VI_STATE(w); \
            if (st == MIS_ELECTED || (st == MIS_HALF && !LSIGN(L_))) eMin = min(eMin, r); \
            else if (st == MIS_UNDECIDED && r < H) { any = true; \
                if (VI_CLASS(w) == CS_CAND) { if (r < u1) { u2 = u1; u1 = r; } else if (r < u2) u2 = r; } } } while (0)
#pragma unroll
        for (u32 k = 0; k < 8; k++) if (k < h.y) MC_SCAN(lv[k], wv[k]);
        for (u32 k = 8; k < h.y; k++) { const u32 ll = l[k]; MC_SCAN(ll, vinfo[LABS(ll)]); }
#undef MC_SCAN
        if (!any) continue;
        const bool big = (int)h.y > maxcsize;
        // ovs[u]: bit 0 = an oversized clause in u's positive list, bit 1 = in its negative list (byte inside an atomically or-ed word)
#define MC_APPLY(L_, W_) do { const u32 u_ = LABS(L_), w_ = (W_); const u32 r_ = VI_RANK(w_); \
            if (VI_STATE(w_) == MIS_UNDECIDED && r_ < H) { \
                if (big) atomicOr((u32*)(ovs + (u_ & ~3u)), (LSIGN(L_) ? 2u : 1u) << (8u * (u_ & 3u))); \
                if (eMin < r_) vinfo[u_] = (w_ & ~7u) | MIS_FROZEN; \
                else { const u32 other = (u1 == r_) ? u2 : u1; if (other < r_) atomicMin(&nbr[u_], other); } } } while (0)
#pragma unroll
        for (u32 k = 0; k < 8; k++) if (k < h.y) MC_APPLY(lv[k], wv[k]);
        for (u32 k = 8; k < h.y; k++) { const u32 ll = l[k]; MC_APPLY(ll, vinfo[LABS(ll)]); }
#undef MC_APPLY
    }
}
__global__ void k_mis_first(const u32* __restrict__ eligible, u32 rBegin, u32 rEnd, u32* vinfo, const u32* __restrict__ nbr,
                            const unsigned char* __restrict__ ovs, u32* __restrict__ blocker, u32* __restrict__ wl, DevCounters* dc, u32 slot,
                            u32* __restrict__ pushList, u32* pushCount) {
    for (u32 r0 = rBegin + blockIdx.x * blockDim.x; r0 < rEnd; r0 += gridDim.x * blockDim.x) {
        const u32 r = r0 + threadIdx.x;
        bool queue = false; u32 v = 0;
        if (r < rEnd) {
            v = eligible[r];
            const u32 w = vinfo[v];
            if (VI_STATE(w) == MIS_UNDECIDED) {
                const u32 m = nbr[v];
                if (m == NOVAR) {
                    if (VI_CLASS(w) == CS_STOP) { vinfo[v] = (w & ~7u) | MIS_LIVESTOP; atomicMin(&dc->misStopRank, r); }
                    else if (ovs[v] & 1) vinfo[v] = (w & ~7u) | MIS_FROZEN;                 // oversized clause in the positive list
                    else if (ovs[v] & 2) { vinfo[v] = (w & ~7u) | MIS_HALF; atomicAdd(&dc->scratch[8], 1u); pushList[atomicAdd(pushCount, 1u)] = v | 0x80000000u; }
                    else { vinfo[v] = (w & ~7u) | MIS_ELECTED; pushList[atomicAdd(pushCount, 1u)] = v; }
                } else { blocker[v] = eligible[m]; queue = true; }
            }
        }
        const u32 mq = __ballot_sync(0xffffffffu, queue);
        if (mq) {
            const u32 leader = __ffs(mq) - 1;
            u32 base = 0;
            if (laneId() == leader) base = atomicAdd(&dc->wlCnt[slot], __popc(mq));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (queue) wl[base + __popc(mq & lanemaskLt())] = v;
        }
    }
}

// withHalf: also the MIS_HALF variables - the walk order of everything that froze something (k_frozen12)
__global__ void k_elect_flags(const u32* __restrict__ eligible, const u32* __restrict__ vinfo, u32 rEnd, u32* __restrict__ flags, int withHalf) {
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < rEnd; r += gridDim.x * blockDim.x) {
        const u32 st = VI_STATE(vinfo[eligible[r]]);
        flags[r] = (st == MIS_ELECTED || (withHalf && st == MIS_HALF)) ? 1u : 0u;
    }
}
__global__ void k_elect_scatter(const u32* __restrict__ eligible, const u32* __restrict__ flags, u32 rEnd,
                                const u32* __restrict__ pos, u32* __restrict__ elected) {
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < rEnd; r += gridDim.x * blockDim.x)
        if (flags[r]) elected[pos[r]] = eligible[r];
}

// ------------------------------------------------------------------ function-table indices
// Walks the elected variables like depFreeze_d (lcve.cu:33-62) - positive list then negative
// list, clauses in clause-index order (the unsorted-list policy), literals in
// clause order - until 12 distinct frozen variables are known.  One CTA; tiny.
__global__ void __launch_bounds__(128) k_frozen12(const u32* __restrict__ elected, const u32* nList, const u32* __restrict__ vinfo,
                                                  DevCounters* dc, const uint4* __restrict__ hdr,
                                                  const u32* __restrict__ pool, const u32* __restrict__ otStart,
                                                  const u32* __restrict__ otSize, const u32* __restrict__ occurs,
                                                  u32* __restrict__ varcore, u32* __restrict__ prev12) {
    __shared__ u32 fv[MAXFUNVAR];
    __shared__ u32 nf, sMin[4], cur;
    const u32 nE = *nList;
    if (threadIdx.x < MAXFUNVAR) { const u32 old = prev12[threadIdx.x]; if (old != NOVAR) varcore[old] = NOVAR; }
    if (threadIdx.x == 0) nf = 0;
    __syncthreads();
    for (u32 ei = 0; ei < nE && nf < MAXFUNVAR; ei++) {
        const u32 x = elected[ei];
        const u32 nsides = VI_STATE(vinfo[x]) == MIS_HALF ? 1u : 2u;   // MIS_HALF: only its positive list froze anything
        for (u32 side = 0; side < nsides && nf < MAXFUNVAR; side++) {
            const u32 lit = V2L(x) | side;
            const u32 n = otSize[lit];
            const u32* list = occurs + otStart[lit];
            u64 last = 0;  // clause index + 1 of the last processed clause
            while (nf < MAXFUNVAR) {
                u32 m = NOVAR;
                for (u32 j = threadIdx.x; j < n; j += blockDim.x) { const u32 ci = list[j]; if ((u64)ci + 1 > last && ci < m) m = ci; }
                for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
                if ((threadIdx.x & 31u) == 0) sMin[threadIdx.x >> 5] = m;
                __syncthreads();
                if (threadIdx.x == 0) {
                    u32 mm = min(min(sMin[0], sMin[1]), min(sMin[2], sMin[3]));
                    cur = mm;
                    if (mm != NOVAR) {
                        const uint4 h = hdr[mm];
                        if (!C_DELETED(h.w)) {
                            const u32* l = pool + h.x;
                            for (u32 k = 0; k < h.y && nf < MAXFUNVAR; k++) {
                                const u32 v = LABS(l[k]);
                                if (v == x) continue;
                                bool seen = false;
                                for (u32 q = 0; q < nf; q++) seen |= fv[q] == v;
                                if (!seen) fv[nf++] = v;
                            }
                        }
                    }
                }
                __syncthreads();
                if (cur == NOVAR) break;
                last = (u64)cur + 1;
                __syncthreads();
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < MAXFUNVAR) {
        const bool on = threadIdx.x < nf;
        prev12[threadIdx.x] = on ? fv[threadIdx.x] : NOVAR;
        if (on) varcore[fv[threadIdx.x]] = threadIdx.x;
    }
    if (threadIdx.x == 0) dc->nFrozen = nf;
}

// ------------------------------------------------------------------ -lcvefast helpers
// mis_oversize (lcve.cu:108-116) for every variable at once: one pass over the clause headers, literals read only of the
// (rare) clauses longer than lcveclausemax
__global__ void k_mark_oversize(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n, int maxcsize, unsigned char* __restrict__ ovs) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w) || (int)h.y <= maxcsize) continue;
        for (u32 k = 0; k < h.y; k++) ovs[LABS(pool[h.x + k])] = 1;
    }
}
// mis_freeze (lcve.cu:133-148): every variable sharing a clause with an elected one is frozen
template <int GS>
__global__ void __launch_bounds__(256) k_mark_frozen(const u32* __restrict__ elected, const u32* __restrict__ count, const uint4* __restrict__ hdr,
                                                     const u32* __restrict__ pool, const u32* __restrict__ otStart, const u32* __restrict__ otSize,
                                                     const u32* __restrict__ occurs, unsigned char* __restrict__ frozen) {
    const u32 n = *count;
    const u32 lane = threadIdx.x & (u32)(GS - 1);
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    for (u32 it = (blockIdx.x * blockDim.x + threadIdx.x) / GS; it < n; it += groupsPerGrid) {
        const u32 v = elected[it];
        for (u32 side = 0; side < 2; side++) {
            const u32 lit = V2L(v) | side;
            const u32 m = otSize[lit];
            const u32* list = occurs + otStart[lit];
            for (u32 j = lane; j < m; j += GS) {
                const uint4 h = hdr[list[j]];
                if (C_DELETED(h.w)) continue;
                for (u32 k = 0; k < h.y; k++) { const u32 u = LABS(pool[h.x + k]); if (u != v) frozen[u] = 1; }
            }
        }
    }
}
// mis_collect_k + mapfrozen_k (lcve.cu:205-215, 253-260): only indices < 12 are observable (function.cuh:35,143); the
// reference hands them out through an atomic counter, here the frozen variables are taken in variable order.  One CTA.
__global__ void __launch_bounds__(1024) k_first12_fast(const unsigned char* __restrict__ frozen, u32 V, DevCounters* dc, u32* __restrict__ varcore,
                                                       u32* __restrict__ prev12) {
    __shared__ u32 fv[MAXFUNVAR];
    __shared__ u32 nf, wcnt[32];
    if (threadIdx.x < MAXFUNVAR) { const u32 old = prev12[threadIdx.x]; if (old != NOVAR) varcore[old] = NOVAR; }
    if (threadIdx.x == 0) nf = 0;
    __syncthreads();
    for (u32 base = 1; base <= V && nf < MAXFUNVAR; base += 1024) {
        const u32 v = base + threadIdx.x;
        const bool f = v <= V && frozen[v];
        const u32 m = __ballot_sync(0xffffffffu, f);
        if ((threadIdx.x & 31u) == 0) wcnt[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        u32 before = nf;
        for (u32 w = 0; w < (threadIdx.x >> 5); w++) before += wcnt[w];
        const u32 slot = before + __popc(m & lanemaskLt());
        if (f && slot < MAXFUNVAR) fv[slot] = v;
        __syncthreads();
        if (threadIdx.x == 0) { u32 t = nf; for (int w = 0; w < 32; w++) t += wcnt[w]; nf = t; }
        __syncthreads();
    }
    const u32 cnt = nf < MAXFUNVAR ? nf : MAXFUNVAR;
    if (threadIdx.x < MAXFUNVAR) {
        const bool on = threadIdx.x < cnt;
        prev12[threadIdx.x] = on ? fv[threadIdx.x] : NOVAR;
        if (on) varcore[fv[threadIdx.x]] = threadIdx.x;
    }
    if (threadIdx.x == 0) dc->nFrozen = cnt;
}

// ------------------------------------------------------------------ host driver
#define MIS_BATCH 6   // MIS rounds queued per host round-trip (an empty round costs ~4 us, a round trip ~40 us)

// kernel profile mode: fetch and clear the device byte counter, attribute it to the kernel launched last
static void profCollect(Ctx* c, int kid) {
    if (!c->ktOn || kid < 0) return;
    unsigned long long b = 0;
    cudaMemcpyAsync(&b, &c->dc->profBytes, 8, cudaMemcpyDeviceToHost, c->stream);
    cudaMemsetAsync(&c->dc->profBytes, 0, 8, c->stream);
    cudaStreamSynchronize(c->stream);
    c->ktBytes[kid] += (double)b;
}

int runLCVE(Ctx* c) {
    const u32 V = c->V;
    unsigned long long* prof = c->ktOn ? (unsigned long long*)&c->dc->profBytes : nullptr;
    if (prof) cudaMemsetAsync(prof, 0, 8, c->stream);
    const u32 pmax = c->o.mu_pos << c->multiplier, nmax = c->o.mu_neg << c->multiplier;
    u32* vinfo = c->rank;
    u32* blocker = c->sortV;   // free once the radix sort is done
    u32* nbr = c->sortK;
    unsigned char* ovs = c->mis;
    CUDA_TRY(cudaMemsetAsync(&c->dc->scratch[6], 0, 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(&c->dc->scratch[8], 0, 4, c->stream));   // number of MIS_HALF decisions
    const int fast = c->o.lcve_fast != 0;
    if (fast) { const int rc0 = syncCounters(c); if (rc0) return rc0; }   // flags bit 3 (a clause with >= 2^14 literals) must be current
    const int maxcsize = fast ? 0x7FFFFFFF : c->o.lcve_clause_max;   // -lcvefast: oversized clauses are filtered up front, never seen by the rounds
    const unsigned char* ovsFast = nullptr;
    if (fast && (c->o.lcve_clause_max < (1 << 14) || (c->hdc->flags & 8u))) {   // flags bit 3: some clause has >= 2^14 literals (k_ot_count)
        CUDA_TRY(cudaMemsetAsync(c->needSort, 0, (size_t)V + 1, c->stream));
        LAUNCH(c, k_mark_oversize, gridFor(c->hdc->numCls, 256), 256, 0, c->hdr[c->cur], c->pool[c->cur], c->hdc->numCls, c->o.lcve_clause_max, c->needSort);
        ovsFast = c->needSort;
    }
    LAUNCH(c, k_scores, gridFor(V, 256), 256, 0, c->hist, c->vstate, c->assumed, V, pmax, nmax, c->o.lcve_max_occurs,
           c->scores, c->eligible, c->cstat, c->dc, fast, ovsFast);
    KB(c, 18.0 * V);
    // the largest score decides how many radix passes the sort needs (usually 2 of 4); the round trip is cheaper than the two
    // spare passes even at V = 100 k (six more launches: cfg1 3.48 -> 3.62 ms without it, profiles/r02_ab_c19.jsonl)
    int rc = syncCounters(c);
    if (rc) return rc;
    u32 bits = 0;
    while (bits < 32 && (c->hdc->scratch[6] >> bits)) bits++;
    radixSortPairs(c, c->scores, c->eligible, c->sortK, c->sortV, V, bits);
    // the per-variable MIS scratch is cleared with coalesced memsets; k_rank only scatters the election words
    CUDA_TRY(cudaMemsetAsync(blocker, 0, (size_t)(V + 1) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(nbr, 0xFF, (size_t)(V + 1) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(ovs, 0, (size_t)V + 1, c->stream));
    LAUNCH(c, k_rank, gridFor(V, 256), 256, 0, c->eligible, c->cstat, V, vinfo, c->dc);
    KB(c, 9.0 * V);
    rc = syncCounters(c);
    if (rc) return rc;
    const u32 firstStop = c->hdc->firstStop < V ? c->hdc->firstStop : V;   // everything ranked before it is walked for sure
    const u32 nCls = c->hdc->numCls;
    // lanes per candidate: variables of Tseitin-like formulas have a handful of short clauses
    const bool smallGroups = c->numLiterals <= (u64)24 * V;
    const u64 avgOcc = c->numLiterals / V + 1;

    // Dense neighbourhoods (uniform k-SAT with many occurrences: hundreds of neighbours per variable):
    // a few thousand elected variables freeze almost everything by push, so the walk starts with a
    // short rank prefix and later chunks only queue what is still undecided.
    const u64 avgK = c->numClauses ? c->numLiterals / c->numClauses + 1 : 2;
    const bool dense = 2 * avgOcc * (avgK > 1 ? avgK - 1 : 1) >= 128;

    u32 hPrev = 0, stopRank = NOVAR, hEnd = 0;
    u32 round = 0;   // parity / counter slot of the next MIS round
    while (hPrev < V && stopRank == NOVAR) {
        u64 h64 = hPrev ? (u64)hPrev * 8 : (firstStop > 8192 ? firstStop : 8192);
        if (dense && !hPrev) { const u64 pre = V / 64 > 8192 ? V / 64 : 8192; if (h64 > pre) h64 = pre; }
        const u32 H = (u32)(h64 > V ? V : h64);
        u32* wlIn = (round & 1u) ? c->wlB : c->wlA;
        u32 n = H - hPrev;   // upper bound of the worklist until the first read-back
        // first round of the chunk: one streaming pass over the clauses when many candidates are
        // undecided, otherwise the candidates walk their own lists
        bool clausePass = (u64)n * avgOcc * 2 > nCls;
        if (dense && hPrev) {
            // the elected variables of the earlier chunks have frozen most of this one (hundreds of neighbours each): what is
            // left walks its own lists - no read-back to count it first, the rounds stride over the device-side count
            LAUNCH(c, k_mis_fill, gridFor(H - hPrev, 256), 256, 0, c->eligible, vinfo, hPrev, H, wlIn, c->dc, round % 3u);
            KB(c, 12.0 * (H - hPrev));
            clausePass = false;
        }
        if (clausePass) {
            LAUNCH(c, k_mis_clauses, gridFor(nCls, 256), 256, 0, c->hdr[c->cur], c->pool[c->cur], nCls, vinfo, nbr, ovs, H, maxcsize);
            KB(c, 16.0 * nCls + 8.0 * c->numLiterals);   // headers, literals, election words
            u32* pushCount = &c->dc->scratch[2];
            CUDA_TRY(cudaMemsetAsync(pushCount, 0, 4, c->stream));
            LAUNCH(c, k_mis_first, gridFor(H - hPrev, 256), 256, 0, c->eligible, hPrev, H, vinfo, nbr, ovs, blocker, wlIn, c->dc, round % 3u,
                   c->flagA, pushCount);
            KB(c, 21.0 * (H - hPrev));
            if (smallGroups)
                LAUNCH(c, k_mis_push<8>, gridFor((u64)(H - hPrev) * 8, 256), 256, 0, c->flagA, pushCount, c->hdr[c->cur], c->pool[c->cur],
                       c->otStart, c->otSize, c->occurs, vinfo, prof);
            else
                LAUNCH(c, k_mis_push<32>, gridFor((u64)(H - hPrev) * 32, 256), 256, 0, c->flagA, pushCount, c->hdr[c->cur], c->pool[c->cur],
                       c->otStart, c->otSize, c->occurs, vinfo, prof);
            profCollect(c, c->ktLastId);
            n = H - hPrev;
        } else if (!(dense && hPrev))
            LAUNCH(c, k_mis_fill, gridFor(H - hPrev, 256), 256, 0, c->eligible, vinfo, hPrev, H, wlIn, c->dc, round % 3u);
            KB(c, 12.0 * (H - hPrev));
        u32 guard = 0;
        while (n) {
            if (++guard > 100000u) { snprintf(c->err, sizeof c->err, "MIS did not converge"); return SIGMA_AWAKEN_FAIL; }
            const u32 gs = smallGroups ? 8u : 32u;
            u64 blocks = ((u64)n * gs + 255) / 256;
            if (blocks > 148ull * 32) blocks = 148ull * 32;
            for (int b = 0; b < MIS_BATCH; b++, round++) {
                // the worklist shrinks fast and the later rounds of a batch are often empty: they stride over it with at most
                // eight CTAs per SM instead of paying for the first round's grid
                const u32 grid = b == 0 ? (u32)blocks : (u32)(blocks < 148ull * 8 ? blocks : 148ull * 8);
                if (smallGroups)
                    LAUNCH(c, k_mis_round<8>, grid, 256, 0, c->wlA, c->wlB, round, c->dc, c->hdr[c->cur], c->pool[c->cur], c->otStart,
                           c->otSize, c->occurs, vinfo, blocker, maxcsize, prof);
                else
                    LAUNCH(c, k_mis_round<32>, grid, 256, 0, c->wlA, c->wlB, round, c->dc, c->hdr[c->cur], c->pool[c->cur], c->otStart,
                           c->otSize, c->occurs, vinfo, blocker, maxcsize, prof);
            }
            profCollect(c, c->ktLastId);
            CUDA_TRY(cudaMemcpyAsync(c->hdc, c->dc, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            n = c->hdc->wlCnt[round % 3u];
        }
        stopRank = c->hdc->misStopRank;
        hPrev = H;
        hEnd = H;
    }
    const u32 rEnd = stopRank < hEnd ? stopRank : hEnd;
    if (rEnd) {
        LAUNCH(c, k_elect_flags, gridFor(rEnd, 256), 256, 0, c->eligible, vinfo, rEnd, c->flagA, 0);
        scanExclusiveU32(c, c->flagA, c->flagB, rEnd, 0, &c->dc->numElected);
        LAUNCH(c, k_elect_scatter, gridFor(rEnd, 256), 256, 0, c->eligible, c->flagA, rEnd, c->flagB, c->elected);
    }
    if (c->o.ve_fun_en && fast) {
        unsigned char* frozen = c->cstat;   // read by k_rank only
        CUDA_TRY(cudaMemsetAsync(frozen, 0, (size_t)V + 1, c->stream));
        if (rEnd) {
            if (smallGroups) LAUNCH(c, k_mark_frozen<8>, gridFor((u64)rEnd * 8, 256), 256, 0, c->elected, &c->dc->numElected, c->hdr[c->cur], c->pool[c->cur],
                                    c->otStart, c->otSize, c->occurs, frozen);
            else LAUNCH(c, k_mark_frozen<32>, gridFor((u64)rEnd * 32, 256), 256, 0, c->elected, &c->dc->numElected, c->hdr[c->cur], c->pool[c->cur],
                        c->otStart, c->otSize, c->occurs, frozen);
        }
        LAUNCH(c, k_first12_fast, 1, 1024, 0, frozen, V, c->dc, c->varcore, c->dc->froz12);
    } else if (c->o.ve_fun_en) {
        const u32* walk = c->elected; const u32* nWalk = &c->dc->numElected;
        if (rEnd && c->hdc->scratch[8]) {   // some variables froze only their positive side: the walk order includes them
            LAUNCH(c, k_elect_flags, gridFor(rEnd, 256), 256, 0, c->eligible, vinfo, rEnd, c->flagA, 1);
            scanExclusiveU32(c, c->flagA, c->flagB, rEnd, 0, &c->dc->scratch[9]);
            LAUNCH(c, k_elect_scatter, gridFor(rEnd, 256), 256, 0, c->eligible, c->flagA, rEnd, c->flagB, blocker);
            walk = blocker; nWalk = &c->dc->scratch[9];
        }
        LAUNCH(c, k_frozen12, 1, 128, 0, walk, nWalk, vinfo, c->dc, c->hdr[c->cur], c->pool[c->cur], c->otStart, c->otSize, c->occurs,
               c->varcore, c->dc->froz12);
    }
    return syncCounters(c);
}
