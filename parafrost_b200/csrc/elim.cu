// parafrost_b200/csrc/elim.cu -- the elimination kernels: SUB (self-subsuming strengthening +
// backward subsumption), BVE (gate detection, bounded resolvent counting, compacted resolvent
// emission), BCE and ERE.
//
// Reference semantics reproduced bit for bit (one THREAD per elected variable there):
//   sub_k       src/gpu/subsume.cuh:402-484
//   ve_k_1/2    src/gpu/bounded.cuh:282-544, resolve.cuh, and.cuh, equivalence.cuh, ifthenelse.cuh,
//               xor.cuh, function.cuh, elimination.cuh, model.cuh
//   bce_k       src/gpu/blocked.cuh:26-97
//   ere_k       src/gpu/redundancy.cuh:99-174
//
// B200 mapping: one WARP per elected variable.  The elected variables are independent (no
// clause contains two of them), so a warp owns every clause it writes.  Regular work - the
// |pos| x |neg| resolvent merges, list scans, subsumption candidates - is strided over the 32
// lanes and combined with ballots / shuffles; the irregular gate searches are executed
// redundantly by all lanes (uniform control flow, broadcast loads) with lane 0 doing the
// writes, each followed by __syncwarp().  Results do not depend on the lane mapping: counts
// are sums, "first match in list order" is taken with ballot + ffs.
#include <cstdlib>

#include "common.cuh"

struct G {
    uint4* hdr; u32* pool;
    const u32* otStart; u32* otSize; u32* occurs;
    const uint4* key;   // OLIST_CMP keys of this round (k_ot_count); lists are sorted by (key, index)
    const u32* bloom; u32 bloomMask;   // ERE: one bit per hashed key of a live clause (k_ere_bloom)
    const u32* elected; unsigned char* eliminated; const u32* vorg; const u32* varcore;
    u32* units; u32 unitsCap; u32* resolved; u32 resolvedCap;
    u32 *veType, *veUcnt, *veRpos; u64* veRref;
    DevCounters* dc;
    KOpts k;
    u32 numElected;
};

// Cooperative group of GS lanes (4, 8 or 32) inside a warp: one group per elected variable.
// Small variables (Tseitin / adder gates: a handful of short clauses per side) run 8 or 4 to a
// warp instead of wasting 28 lanes; the classes are built by k_bin_elected.
template <int GS> struct GT : G {};
#define LANE (threadIdx.x & (u32)(GS - 1))
#define GBASE ((threadIdx.x & 31u) & ~(u32)(GS - 1))
#define FULL (GS == 32 ? 0xffffffffu : (((1u << (GS & 31)) - 1u) << GBASE))
#define BALLOT(p) (__ballot_sync(FULL, (p)) >> GBASE)
#define LTMASK ((1u << LANE) - 1u)
#define GSYNC() __syncwarp(FULL)
template <int GS> __device__ __forceinline__ u32 gSum(u32 v) {
#pragma unroll
    for (int o = GS / 2; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o, GS);
    return v;
}
template <int GS> __device__ __forceinline__ u32 gIncl(u32 v) {
#pragma unroll
    for (int o = 1; o < GS; o <<= 1) { const u32 t = __shfl_up_sync(FULL, v, o, GS); if (LANE >= (u32)o) v += t; }
    return v;
}

// ------------------------------------------------------------------ small helpers
template <int GS> __device__ __forceinline__ void setBits(GT<GS>& g, u32 ci, u32 set, u32 clr) {
    if (LANE == 0) { const u32 w = g.hdr[ci].w; g.hdr[ci].w = (w & ~clr) | set; }
    GSYNC();
}
template <int GS> __device__ __forceinline__ void melt(GT<GS>& g, u32 ci) { setBits(g, ci, CB_MOLTEN, 0); }
template <int GS> __device__ __forceinline__ void freeze(GT<GS>& g, u32 ci) { setBits(g, ci, 0, CB_MOLTEN); }
template <int GS> __device__ __forceinline__ void markDeleted(GT<GS>& g, u32 ci) { setBits(g, ci, CB_DELETED, CB_ST_MASK); }

// Pull the clauses of a small variable into L1 up front: the gate searches below chase
// list entry -> header -> literals serially, so without this every step is an L2/HBM round trip.
template <int GS> __device__ __forceinline__ void prefetchLists(GT<GS>& g, const u32* P, u32 np, const u32* N, u32 nn) {
    if (np + nn > 512u) return;   // up to 16 clauses per lane of a full warp (uniform k-SAT: ~50 occurrences per literal)
    for (u32 j = LANE; j < np + nn; j += GS) {
        const u32 ci = j < np ? P[j] : N[j - np];
        const uint4 h = g.hdr[ci];
        const u32* l = g.pool + h.x;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(l));
        if (h.y > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(l + h.y - 1));
    }
}

// resolvent length on x, 0 if tautology (elimination.cuh:180-205)
__device__ __forceinline__ int mergeLen(const u32* __restrict__ a, int n1, const u32* __restrict__ b, int n2, u32 x) {
    int it1 = 0, it2 = 0, len = n1 + n2 - 2;
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = LABS(lit1), v2 = LABS(lit2);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (IS_TAUT(lit1, lit2)) return 0;
        else if (v1 < v2) it1++;
        else if (v2 < v1) it2++;
        else { it1++; it2++; len--; }
    }
    return len;
}
// resolvent into out (elimination.cuh:277-310); returns length, 0 if tautology; sig accumulated
__device__ __forceinline__ int mergeOut(const u32* a, int n1, const u32* b, int n2, u32 x, u32* out, u32& sig) {
    int it1 = 0, it2 = 0, len = 0;
    sig = 0;
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = LABS(lit1), v2 = LABS(lit2);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (IS_TAUT(lit1, lit2)) return 0;
        else if (v1 < v2) { it1++; out[len++] = lit1; sig |= MAPHASH(lit1); }
        else if (v2 < v1) { it2++; out[len++] = lit2; sig |= MAPHASH(lit2); }
        else { it1++; it2++; out[len++] = lit1; sig |= MAPHASH(lit1); }
    }
    while (it1 < n1) { const u32 l = a[it1++]; if (LABS(l) != x) { out[len++] = l; sig |= MAPHASH(l); } }
    while (it2 < n2) { const u32 l = b[it2++]; if (LABS(l) != x) { out[len++] = l; sig |= MAPHASH(l); } }
    return len;
}
__device__ __forceinline__ bool isTautology(const u32* a, int n1, const u32* b, int n2, u32 x) {
    int it1 = 0, it2 = 0;
    while (it1 < n1 && it2 < n2) {
        const u32 v1 = LABS(a[it1]), v2 = LABS(b[it2]);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (IS_TAUT(a[it1], b[it2])) return true;
        else if (v1 < v2) it1++;
        else if (v2 < v1) it2++;
        else { it1++; it2++; }
    }
    return false;
}
__device__ __forceinline__ u32 sigOf(const u32* l, int n) {
    u32 s = 0;
    for (int k = 0; k < n; k++) s |= MAPHASH(l[k]);
    return s;
}

// ------------------------------------------------------------------ witness stack (model.cuh:29-53)
// lane 0 writes; callers reserve with atomicAdd on dc->resolvedSize
template <int GS> __device__ __forceinline__ void saveWitness(GT<GS>& g, u32*& saved, u32 witness) {
    *saved++ = V2L(g.vorg[LABS(witness)]) | LSIGN(witness);
    *saved++ = 1;
}
template <int GS> __device__ __forceinline__ void saveClause(GT<GS>& g, u32*& saved, const uint4 h, u32 witlit) {
    u32* first = saved; u32* wit = saved;
    const u32* l = g.pool + h.x;
    for (u32 k = 0; k < h.y; k++) {
        const u32 lit = l[k];
        if (lit == witlit) wit = saved;
        *saved++ = V2L(g.vorg[LABS(lit)]) | LSIGN(lit);
    }
    const u32 t = *first; *first = *wit; *wit = t;
    *saved++ = h.y;
}
// reserve n words of the witness stack; returns NULL on overflow (flagged; the reference only asserts)
template <int GS> __device__ __forceinline__ u32* jumpResolved(GT<GS>& g, u32 n) {
    u32 base = 0;
    if (LANE == 0) base = atomicAdd(&g.dc->resolvedSize, n);
    base = __shfl_sync(FULL, base, 0, GS);
    if ((u64)base + n > g.resolvedCap) { if (LANE == 0) atomicOr(&g.dc->flags, 1u); return nullptr; }
    return g.resolved + base;
}
// count originals / their literals of a list (all lanes redundantly; lists are short)
template <int GS> __device__ __forceinline__ void countOrgsLits(GT<GS>& g, const u32* list, u32 n, u32& cls, u32& lits) {
    u32 c = 0, l = 0;
    for (u32 j = LANE; j < n; j += GS) { const uint4 h = g.hdr[list[j]]; if (C_ORIGINAL(h.w)) c++, l += h.y; }
    cls = gSum<GS>(c); lits = gSum<GS>(l);
}
// save the originals of `list` (witness literal first) followed by the witness unit of `wit`
// `reserveCls`: clause count used for the reservation (the reference reserves pOrgs/nOrgs words
//  in phase 1, elimination.cuh:455-476, which equals the number of originals except when learnt
//  clauses sit in the list of a first-call run; reserve what the reference reserves)
template <int GS> __device__ __forceinline__ void saveSide(GT<GS>& g, const u32* list, u32 n, u32 witlit, u32 wit, u32 reserveCls) {
    u32 cls, lits;
    countOrgsLits(g, list, n, cls, lits);
    u32* saved = jumpResolved(g, reserveCls + lits + 2);
    if (saved && LANE == 0) {
        u32* end = saved + reserveCls + lits + 2;
        for (u32 j = 0; j < n; j++) { const uint4 h = g.hdr[list[j]]; if (C_ORIGINAL(h.w)) saveClause(g, saved, h, witlit); }
        saveWitness(g, saved, wit);
        while (saved < end) *saved++ = 0;  // unreachable unless reserveCls > #originals
    }
    GSYNC();
}
template <int GS> __device__ __forceinline__ void deleteAll(GT<GS>& g, const u32* list, u32 n) {
    for (u32 j = LANE; j < n; j += GS) { const u32 ci = list[j]; const u32 w = g.hdr[ci].w; g.hdr[ci].w = (w & ~CB_ST_MASK) | CB_DELETED; }
    GSYNC();
}
// toblivion with witness saving (elimination.cuh:443-492)
template <int GS> __device__ void toblivionSave(GT<GS>& g, u32 p, u32 n, u32 pOrgs, u32 nOrgs, const u32* P, u32 np, const u32* N, u32 nn) {
    if (pOrgs > nOrgs) saveSide(g, N, nn, n, p, nOrgs);
    else saveSide(g, P, np, p, n, pOrgs);
    deleteAll(g, P, np);
    deleteAll(g, N, nn);
    if (LANE == 0) { g.otSize[p] = 0; g.otSize[n] = 0; }
    GSYNC();
}
template <int GS> __device__ __forceinline__ void freezeBinaries(GT<GS>& g, const u32* list, u32 n) {
    for (u32 j = LANE; j < n; j += GS) { const u32 ci = list[j]; const uint4 h = g.hdr[ci]; if (C_ORIGINAL(h.w) && h.y == 2) g.hdr[ci].w = h.w & ~CB_MOLTEN; }
    GSYNC();
}
template <int GS> __device__ __forceinline__ void freezeClauses(GT<GS>& g, const u32* P, u32 np, const u32* N, u32 nn) {
    for (u32 j = LANE; j < np; j += GS) { const u32 ci = P[j]; const u32 w = g.hdr[ci].w; if (C_ORIGINAL(w) && C_MOLTEN(w)) g.hdr[ci].w = w & ~CB_MOLTEN; }
    for (u32 j = LANE; j < nn; j += GS) { const u32 ci = N[j]; const u32 w = g.hdr[ci].w; if (C_ORIGINAL(w) && C_MOLTEN(w)) g.hdr[ci].w = w & ~CB_MOLTEN; }
    GSYNC();
}
template <int GS> __device__ __forceinline__ void freezeArities(GT<GS>& g, const u32* P, u32 np, const u32* N, u32 nn) {
    for (u32 j = LANE; j < np; j += GS) { const u32 ci = P[j]; const uint4 h = g.hdr[ci]; if (h.y > 2 && C_MOLTEN(h.w)) g.hdr[ci].w = h.w & ~CB_MOLTEN; }
    for (u32 j = LANE; j < nn; j += GS) { const u32 ci = N[j]; const uint4 h = g.hdr[ci]; if (h.y > 2 && C_MOLTEN(h.w)) g.hdr[ci].w = h.w & ~CB_MOLTEN; }
    GSYNC();
}
// append the unit clauses of a list to the units vector, list order (elimination.cuh:95-105)
template <int GS> __device__ void appendUnits(GT<GS>& g, const u32* list, u32 n, u32& cursor) {
    for (u32 base = 0; base < n; base += GS) {
        const u32 j = base + LANE;
        u32 lit = 0; bool is = false;
        if (j < n) { const uint4 h = g.hdr[list[j]]; if (h.y == 1) { is = true; lit = g.pool[h.x]; } }
        const u32 m = BALLOT(is);
        if (is) { const u32 slot = cursor + __popc(m & LTMASK); if (slot < g.unitsCap) g.units[slot] = lit; else atomicOr(&g.dc->flags, 2u); }
        cursor += __popc(m);
    }
}
template <int GS> __device__ __forceinline__ u32 reserveUnits(GT<GS>& g, u32 n) {
    u32 base = 0;
    if (LANE == 0) base = atomicAdd(&g.dc->numUnits, n);
    return __shfl_sync(FULL, base, 0, GS);
}

// ------------------------------------------------------------------ pair counting
// mode 0: countResolvents without the clause bound (resolve.cuh:28-109)
// mode 1: countResolvents bounded (resolve.cuh:111-199)
// mode 2: countSubstituted (elimination.cuh:365-441)     pairs with different molten flags
// mode 3: countCoreSubstituted (function.cuh:181-257)    pairs not both molten
// returns true when the elimination is NOT possible (bound / size / packing limits exceeded)
template <int GS> __device__ bool countPairs(GT<GS>& g, int mode, u32 x, const u32* M, u32 nm, const u32* O, u32 no, u32 nClsBefore,
                           u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    const u32 rlimit = g.k.ve_clause_max;
    u32 units = 0, cls = 0, lits = 0;
    bool big = false;
    const u64 total = (u64)nm * no;
    u32 iter = 0;
    bool fail = false;
    for (u64 t0 = 0; t0 < total; t0 += GS, iter++) {
        const u64 t = t0 + LANE;
        if (t < total) {
            const u32 i = (u32)(t / no), j = (u32)(t - (u64)i * no);
            const uint4 hi = g.hdr[M[i]];
            if (!C_LEARNT(hi.w)) {
                const uint4 hj = g.hdr[O[j]];
                bool take;
                if (mode <= 1) take = !C_LEARNT(hj.w);
                else if (mode == 2) take = C_ORIGINAL(hj.w) && ((C_MOLTEN(hi.w) != 0) != (C_MOLTEN(hj.w) != 0));
                else take = C_ORIGINAL(hj.w) && (!C_MOLTEN(hi.w) || !C_MOLTEN(hj.w));
                if (take) {
                    const int rsize = mergeLen(g.pool + hi.x, (int)hi.y, g.pool + hj.x, (int)hj.y, x);
                    if (rsize == 1) units++;
                    else if (rsize) { cls++; lits += (u32)rsize; if (rlimit && (u32)rsize > rlimit) big = true; }
                }
            }
        }
        if ((iter & 7u) == 7u) {  // early exit like the serial loop's `return`
            if (__any_sync(FULL, big)) { fail = true; break; }
            if (mode && gSum<GS>(cls) > nClsBefore) { fail = true; break; }
        }
    }
    nElements = gSum<GS>(units); nAddedCls = gSum<GS>(cls); nAddedLits = gSum<GS>(lits);
    if (__any_sync(FULL, big)) fail = true;
    if (mode && nAddedCls > nClsBefore) fail = true;
    if (fail) { if (!nAddedCls) nAddedCls = 1; return true; }
    if (nAddedCls > ADDEDCLS_MAX || nAddedLits > ADDEDLITS_MAX) return true;
    if (mode && g.k.ve_lbound_en) {
        u32 c1, l1, c2, l2;
        countOrgsLits(g, M, nm, c1, l1);
        countOrgsLits(g, O, no, c2, l2);
        if (nAddedLits > l1 + l2) return true;
    }
    return false;
}

// ------------------------------------------------------------------ gates (executed redundantly by all lanes)
// equivalence.cuh:111-171
template <int GS> __device__ u32 findEquGate(GT<GS>& g, u32 p, u32 n, const u32* P, u32 np, const u32* N, u32 nn) {
    if (g.hdr[P[0]].y > 2 || g.hdr[N[0]].y > 2) return 0;
    // find_sfanin
    u32 imp = 0; int nImps = 0; bool multi = false;
    for (u32 j = 0; j < np; j++) {
        const u32 ci = P[j]; const uint4 h = g.hdr[ci];
        if (C_ORIGINAL(h.w) && h.y == 2) {
            imp = LFLIP(g.pool[h.x] ^ g.pool[h.x + 1] ^ p);
            melt(g, ci);
            nImps++;
        }
        if (nImps > 1) { multi = true; break; }
    }
    u32 first = multi ? 0 : imp;
    if (first) {
        u32 second = n; const u32 def = first;
        if (second < first) { first = second; second = def; }
        for (u32 j = 0; j < nn; j++) {
            const u32 ci = N[j]; const uint4 h = g.hdr[ci];
            if (C_ORIGINAL(h.w) && h.y == 2 && g.pool[h.x] == first && g.pool[h.x + 1] == second) { melt(g, ci); return def; }
        }
    }
    freezeBinaries(g, P, np);
    return 0;
}

template <int GS> __device__ __forceinline__ bool clauseHas(GT<GS>& g, const uint4 h, u32 lit) {
    const u32* l = g.pool + h.x;
    for (u32 k = 0; k < h.y; k++) if (l[k] == lit) return true;
    return false;
}
// substitute_single on one clause (equivalence.cuh:28-57); lane 0 only
template <int GS> __device__ void substituteClause(GT<GS>& g, u32 ci, u32 dx, u32 def, u32& nUnits) {
    uint4 h = g.hdr[ci];
    u32* l = g.pool + h.x;
    if (LANE == 0) {
        u32 n = 0;
        for (u32 k = 0; k < h.y; k++) { const u32 lit = l[k]; if (lit == dx) l[n++] = def; else if (lit != def) l[n++] = lit; }
        h.y = n;
        if (n > 1) {
            for (u32 a = 1; a < n; a++) { const u32 t = l[a]; int b = (int)a; for (; b > 0 && t < l[b - 1]; b--) l[b] = l[b - 1]; l[b] = t; }
            h.z = sigOf(l, (int)n);
        }
        g.hdr[ci] = h;
    }
    GSYNC();
    if (g.hdr[ci].y == 1) nUnits++;
}
// equivalence.cuh:59-109
template <int GS> __device__ void substituteSingle(GT<GS>& g, u32 p, u32 n, u32 def, const u32* P, u32 np, const u32* N, u32 nn) {
    const u32 def_f = LFLIP(def);
    u32 nNegUnits = 0, nPosUnits = 0;
    for (u32 j = 0; j < nn; j++) {
        const u32 ci = N[j]; const uint4 h = g.hdr[ci];
        if (C_LEARNT(h.w) || C_MOLTEN(h.w) || clauseHas(g, h, def)) markDeleted(g, ci);
        else substituteClause(g, ci, n, def_f, nNegUnits);
    }
    for (u32 j = 0; j < np; j++) {
        const u32 ci = P[j]; const uint4 h = g.hdr[ci];
        if (C_LEARNT(h.w) || C_MOLTEN(h.w) || clauseHas(g, h, def_f)) markDeleted(g, ci);
        else substituteClause(g, ci, p, def, nPosUnits);
    }
    if (nPosUnits || nNegUnits) {
        // appendUnits appends every size-1 clause of the list (elimination.cuh:95-105)
        u32 cn = 0, cp = 0;
        if (nNegUnits) for (u32 j = 0; j < nn; j++) cn += g.hdr[N[j]].y == 1;
        if (nPosUnits) for (u32 j = 0; j < np; j++) cp += g.hdr[P[j]].y == 1;
        u32 cursor = reserveUnits(g, cn + cp);
        if (nNegUnits) appendUnits(g, N, nn, cursor);
        if (nPosUnits) appendUnits(g, P, np, cursor);
    }
}

// and.cuh:28-113 ; out_c lives in this warp's shared slice
template <int GS> __device__ bool findAOGate(GT<GS>& g, u32 dx, const u32* D, u32 nd, u32 fx, const u32* F, u32 nf, u32 nOrgCls, u32* out_c,
                           u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    if (g.hdr[D[0]].y > 2 || g.hdr[F[nf - 1]].y < 3) return false;
    u32 sig = 0; u32 nImps = 0;
    for (u32 j = 0; j < nd; j++) {
        const u32 ci = D[j]; const uint4 h = g.hdr[ci];
        if (C_ORIGINAL(h.w) && h.y == 2) {
            const u32 imp = LFLIP(g.pool[h.x] ^ g.pool[h.x + 1] ^ dx);
            if (LANE == 0) out_c[nImps] = imp;
            nImps++;
            sig |= MAPHASH(imp);
            melt(g, ci);
        }
    }
    if (nImps > 1) {
        const u32 x = LABS(dx);
        if (LANE == 0) {
            out_c[nImps] = fx;
            for (u32 a = 1; a <= nImps; a++) { const u32 t = out_c[a]; int b = (int)a; for (; b > 0 && t < out_c[b - 1]; b--) out_c[b] = out_c[b - 1]; out_c[b] = t; }
        }
        nImps++;
        sig |= MAPHASH(fx);
        GSYNC();
        for (u32 j = 0; j < nf; j++) {
            const u32 ci = F[j]; const uint4 h = g.hdr[ci];
            if (C_ORIGINAL(h.w) && h.y == nImps && SUBSIG(h.z, sig)) {
                bool eq = true;
                for (u32 k = 0; k < nImps; k++) if (g.pool[h.x + k] != out_c[k]) { eq = false; break; }
                if (!eq) continue;
                melt(g, ci);
                nElements = 0; nAddedCls = 0; nAddedLits = 0;
                if (countPairs(g, 2, x, D, nd, F, nf, nOrgCls, nElements, nAddedCls, nAddedLits)) { freeze(g, ci); break; }
                return true;
            }
        }
    }
    freezeBinaries(g, D, nd);
    return false;
}

// ifthenelse.cuh:28-50 ; returns clause index or NOVAR
template <int GS> __device__ u32 fastEqualityCheck(GT<GS>& g, u32 x, u32 y, u32 z) {
    u32 t;
    if (g.otSize[y] > g.otSize[z]) { t = y; y = z; z = t; }
    if (g.otSize[x] > g.otSize[y]) { t = x; x = y; y = t; }
    const u32 n = g.otSize[x];
    const u32* list = g.occurs + g.otStart[x];
    // sort3
    if (y > z) { t = y; y = z; z = t; }
    if (x > z) { t = x; x = z; z = t; }
    if (x > y) { t = x; x = y; y = t; }
    // The reference returns the first match in (sorted) list order.  All matches are copies of the
    // same clause, i.e. they share the sort key and are ordered by clause index: the first match is
    // the one with the smallest index, which does not need this (foreign) list to be sorted.
    u32 found = NOVAR;
    for (u32 j = LANE; j < n; j += GS) {
        const u32 ci = list[j];
        const uint4 h = g.hdr[ci];
        if (!C_MOLTEN(h.w) && C_ORIGINAL(h.w) && h.y == 3) {
            const u32* l = g.pool + h.x;
            if (l[0] == x && l[1] == y && l[2] == z) found = min(found, ci);
        }
    }
#pragma unroll
    for (int o = GS / 2; o; o >>= 1) found = min(found, __shfl_xor_sync(FULL, found, o, GS));
    return found;
}
// The same search when the variable's clauses sit in a local (shared-memory) copy: a clause {fx, y, z} contains fx, so it
// is in F - the list of fx itself - whichever of the three lists the reference walks; copies of one clause share the
// sort key and follow each other in index order, so the first match in F is the clause the reference finds.
template <int GS> __device__ u32 fastEqualityLocal(GT<GS>& g, const u32* F, u32 nf, u32 x, u32 y, u32 z) {
    u32 t;
    if (y > z) { t = y; y = z; z = t; }
    if (x > z) { t = x; x = z; z = t; }
    if (x > y) { t = x; x = y; y = t; }
    u32 found = NOVAR;
    for (u32 j = LANE; j < nf; j += GS) {
        const u32 ci = F[j];
        const uint4 h = g.hdr[ci];
        if (!C_MOLTEN(h.w) && C_ORIGINAL(h.w) && h.y == 3) {
            const u32* l = g.pool + h.x;
            if (l[0] == x && l[1] == y && l[2] == z) found = min(found, ci);
        }
    }
#pragma unroll
    for (int o = GS / 2; o; o >>= 1) found = min(found, __shfl_xor_sync(FULL, found, o, GS));
    return found;
}
// ifthenelse.cuh:52-125
template <int GS, bool LOCAL = false> __device__ bool findITEGate(GT<GS>& g, u32 dx, const u32* D, u32 nd, u32 fx, const u32* F, u32 nf, u32 nOrgCls,
                            u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    if (g.hdr[D[nd - 1]].y == 2) return false;
    const u32 v = LABS(dx);
    for (u32 i = 0; i < nd; i++) {
        const u32 cii = D[i]; const uint4 hi = g.hdr[cii];
        if (!(C_ORIGINAL(hi.w) && hi.y == 3)) continue;
        u32 xi = g.pool[hi.x], yi = g.pool[hi.x + 1], zi = g.pool[hi.x + 2], t;
        if (yi == dx) { t = xi; xi = yi; yi = t; }
        if (zi == dx) { t = xi; xi = zi; zi = t; }
        for (u32 j = i + 1; j < nd; j++) {
            const u32 cjj = D[j]; const uint4 hj = g.hdr[cjj];
            if (!(C_ORIGINAL(hj.w) && hj.y == 3)) continue;
            u32 xj = g.pool[hj.x], yj = g.pool[hj.x + 1], zj = g.pool[hj.x + 2];
            if (yj == dx) { t = xj; xj = yj; yj = t; }
            if (zj == dx) { t = xj; xj = zj; zj = t; }
            if (LABS(yi) == LABS(zj)) { t = yj; yj = zj; zj = t; }
            if (LABS(zi) == LABS(zj)) continue;
            if (yi != LFLIP(yj)) continue;
            const u32 r1 = LOCAL ? fastEqualityLocal(g, F, nf, fx, yi, LFLIP(zi)) : fastEqualityCheck(g, fx, yi, LFLIP(zi));
            if (r1 == NOVAR) continue;
            const u32 r2 = LOCAL ? fastEqualityLocal(g, F, nf, fx, yj, LFLIP(zj)) : fastEqualityCheck(g, fx, yj, LFLIP(zj));
            if (r2 == NOVAR) continue;
            melt(g, cii); melt(g, cjj); melt(g, r1); melt(g, r2);
            nElements = 0; nAddedCls = 0; nAddedLits = 0;
            if (countPairs(g, 2, v, D, nd, F, nf, nOrgCls, nElements, nAddedCls, nAddedLits)) {
                freeze(g, cii); freeze(g, cjj); freeze(g, r1); freeze(g, r2);
                return false;
            }
            return true;
        }
    }
    return false;
}

// xor.cuh:58-109 ; literals[] in this warp's shared slice (lane 0 writes)
template <int GS> __device__ bool makeArity(GT<GS>& g, u32& parity, u32* literals, int size) {
    const u32 oldparity = parity;
    while (__popc(++parity) & 1) {}
    if (LANE == 0)
        for (int k = 0; k < size; k++) { const u32 bit = k < 32 ? 1u << k : 0u; if ((parity & bit) != (oldparity & bit)) literals[k] = LFLIP(literals[k]); }   // xor.cuh:78: (1UL << k) truncated
    GSYNC();
    u32 best = literals[0];
    u32 minsize = g.otSize[best];
    for (int k = 1; k < size; k++) { const u32 lit = literals[k]; const u32 ls = g.otSize[lit]; if (ls < minsize) { minsize = ls; best = lit; } }
    const u32* list = g.occurs + g.otStart[best];
    // first match in sorted order == the matching clause with the smallest index (see fastEqualityCheck)
    u32 found = NOVAR;
    for (u32 j = LANE; j < minsize; j += GS) {
        const u32 ci = list[j];
        const uint4 h = g.hdr[ci];
        if (C_ORIGINAL(h.w) && (int)h.y == size) {
            bool ok = true;
            const u32* l = g.pool + h.x;
            for (int a = 0; a < size && ok; a++) {  // checkArity
                bool f = false;
                for (int b = 0; b < size; b++) if (l[a] == literals[b]) { f = true; break; }
                ok = f;
            }
            if (ok) found = min(found, ci);
        }
    }
#pragma unroll
    for (int o = GS / 2; o; o >>= 1) found = min(found, __shfl_xor_sync(FULL, found, o, GS));
    if (found != NOVAR) { melt(g, found); return true; }
    return false;
}
// makeArity on the local copy: the flipped clause still holds dx or fx (an even number of literals is flipped), so it is
// looked up in D or in F instead of the shortest list among its literals; same clause found (see fastEqualityLocal).
template <int GS> __device__ bool makeArityLocal(GT<GS>& g, u32& parity, u32* literals, int size, u32 dx, const u32* D, u32 nd, const u32* F, u32 nf) {
    const u32 oldparity = parity;
    while (__popc(++parity) & 1) {}
    if (LANE == 0)
        for (int k = 0; k < size; k++) { const u32 bit = k < 32 ? 1u << k : 0u; if ((parity & bit) != (oldparity & bit)) literals[k] = LFLIP(literals[k]); }
    GSYNC();
    bool hasDx = false;
    for (int k = 0; k < size; k++) hasDx |= literals[k] == dx;
    const u32* list = hasDx ? D : F;
    const u32 n = hasDx ? nd : nf;
    u32 found = NOVAR;
    for (u32 j = LANE; j < n; j += GS) {
        const u32 ci = list[j];
        const uint4 h = g.hdr[ci];
        if (C_ORIGINAL(h.w) && (int)h.y == size) {
            bool ok = true;
            const u32* l = g.pool + h.x;
            for (int a = 0; a < size && ok; a++) {
                bool f = false;
                for (int b = 0; b < size; b++) if (l[a] == literals[b]) { f = true; break; }
                ok = f;
            }
            if (ok) found = min(found, ci);
        }
    }
#pragma unroll
    for (int o = GS / 2; o; o >>= 1) found = min(found, __shfl_xor_sync(FULL, found, o, GS));
    if (found != NOVAR) { melt(g, found); return true; }
    return false;
}
// xor.cuh:111-185
template <int GS, bool LOCAL = false> __device__ bool findXORGate(GT<GS>& g, u32 dx, const u32* D, u32 nd, u32 fx, const u32* F, u32 nf, u32 nOrgCls, u32* out_c,
                            u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    if (g.hdr[D[nd - 1]].y == 2 || g.hdr[F[nf - 1]].y == 2) return false;
    const int maxarity = (int)g.k.xor_max_arity;
    if ((int)g.hdr[D[0]].y - 1 > maxarity) return false;
    const u32 v = LABS(dx);
    for (u32 i = 0; i < nd; i++) {
        const u32 ci = D[i]; const uint4 h = g.hdr[ci];
        if (!C_ORIGINAL(h.w)) continue;
        const int size = (int)h.y, arity = size - 1;
        if (size < 3 || arity > maxarity) continue;
        GSYNC();
        if (LANE == 0) for (int k = 0; k < size; k++) out_c[k] = g.pool[h.x + k];
        GSYNC();
        u32 parity = 0;
        int itargets = arity >= 32 ? 0 : (int)(1u << arity);   // what `1 << arity` yields on the device (shl clamps); explicit, no UB
        while (--itargets && (LOCAL ? makeArityLocal(g, parity, out_c, size, dx, D, nd, F, nf) : makeArity(g, parity, out_c, size))) {}
        if (itargets) freezeArities(g, D, nd, F, nf);
        else {
            melt(g, ci);
            nElements = 0; nAddedCls = 0; nAddedLits = 0;
            if (countPairs(g, 2, v, D, nd, F, nf, nOrgCls, nElements, nAddedCls, nAddedLits)) { freezeArities(g, D, nd, F, nf); break; }
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ function tables (function.cuh)
// One 64x64-bit table = 4096 bits; lane l owns words l and l+32.
__constant__ u64 MAGICCONSTS[6] = {0xaaaaaaaaaaaaaaaaULL, 0xccccccccccccccccULL, 0xf0f0f0f0f0f0f0f0ULL,
                                   0xff00ff00ff00ff00ULL, 0xffff0000ffff0000ULL, 0xffffffff00000000ULL};
struct Fun2 { u64 a, b; };  // words LANE and LANE+32
__device__ __forceinline__ u64 funWord(int v, bool sign, u32 i) {  // clause2fun contribution to word i (function.cuh:86-112)
    if (v < 6) { u64 val = MAGICCONSTS[v]; return sign ? ~val : val; }
    const u32 sv = 1u << (v - 6);
    const bool flip = (i / sv) & 1u;    // val starts as (sign ? ones : 0) and toggles every sv words
    const bool ones = sign != flip;
    return ones ? ~0ULL : 0ULL;
}
// OR of the literals of clause h other than lit, as a table; false if a variable index >= 12
template <int GS> __device__ __forceinline__ bool clauseFun(GT<GS>& g, const uint4 h, u32 lit, Fun2& cls) {
    cls.a = cls.b = 0;
    const u32* l = g.pool + h.x;
    for (u32 k = 0; k < h.y; k++) {
        const u32 other = l[k];
        if (other == lit) continue;
        const u32 mvar = g.varcore[LABS(other)];
        if (mvar >= MAXFUNVAR) return false;
        cls.a |= funWord((int)mvar, LSIGN(other), LANE);
        cls.b |= funWord((int)mvar, LSIGN(other), LANE + 32);
    }
    return true;
}
// true iff buildFunAll(lit) would succeed: every non-learnt clause of the list only has mapped neighbours
template <int GS> __device__ bool funPossible(GT<GS>& g, u32 lit, const u32* list, u32 n) {
    bool bad = false;
    for (u32 j = LANE; j < n; j += GS) {
        const uint4 h = g.hdr[list[j]];
        if (C_LEARNT(h.w)) continue;
        const u32* l = g.pool + h.x;
        for (u32 k = 0; k < h.y; k++) { const u32 other = l[k]; if (other != lit && g.varcore[LABS(other)] >= MAXFUNVAR) { bad = true; break; } }
    }
    return !__any_sync(FULL, bad);
}
template <int GS> __device__ bool buildFunAll(GT<GS>& g, u32 lit, Fun2& f) {  // function.cuh:114-148
    f.a = f.b = ~0ULL;
    const u32 n = g.otSize[lit];
    const u32* list = g.occurs + g.otStart[lit];
    for (u32 j = 0; j < n; j++) {
        const uint4 h = g.hdr[list[j]];
        if (C_LEARNT(h.w)) continue;
        Fun2 cls;
        if (!clauseFun(g, h, lit, cls)) return false;
        f.a &= cls.a; f.b &= cls.b;
    }
    return true;
}
template <int GS> __device__ void buildFunTail(GT<GS>& g, u32 lit, u32 tail, const u32* list, Fun2& fun, bool& core) {  // function.cuh:150-179
    for (u32 j = 0; j < tail; j++) {
        const uint4 h = g.hdr[list[j]];
        if (C_LEARNT(h.w)) continue;
        Fun2 cls;
        clauseFun(g, h, lit, cls);
        fun.a &= cls.a; fun.b &= cls.b;
    }
    if (!__any_sync(FULL, (fun.a | fun.b) != 0)) { melt(g, list[tail]); core = true; }
}
template <int GS> __device__ bool findFunGate(GT<GS>& g, u32 p, u32 n, u32 nOrgCls, const u32* P, u32 np, const u32* N, u32 nn,
                            u32& nElements, u32& nAddedCls, u32& nAddedLits) {  // function.cuh:275-327
    Fun2 pos, neg;
    if (buildFunAll(g, p, pos) && buildFunAll(g, n, neg)) {
        if (!__any_sync(FULL, ((pos.a & neg.a) | (pos.b & neg.b)) != 0)) {
            bool core = false;
            Fun2 fun;
            for (int i = (int)np - 1; i >= 0; i--) {
                fun = neg;
                if (C_ORIGINAL(g.hdr[P[i]].w)) buildFunTail(g, p, (u32)i, P, fun, core);
            }
            for (int i = (int)nn - 1; i >= 0; i--) {
                fun.a = fun.b = ~0ULL;
                if (C_ORIGINAL(g.hdr[N[i]].w)) buildFunTail(g, n, (u32)i, N, fun, core);
            }
            nElements = 0; nAddedCls = 0; nAddedLits = 0;
            if (countPairs(g, 3, LABS(p), P, np, N, nn, nOrgCls, nElements, nAddedCls, nAddedLits)) {
                if (core) freezeClauses(g, P, np, N, nn);
                return false;
            }
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ BVE phase 1 (bounded.cuh:282-394)
#define VE_SLICE 256       // words of shared memory per 32-lane group (out_c of the gate searches, >= SH_MAX_BVE_OUT1)
#define VE_SLICE_SMALL 64  // per 4/8-lane group: those classes hold variables with <= BIN_T8 occurrences
#define BIN_T4 12u         // occurrences (pos + neg) up to which a variable runs on a 4-lane group
#define BIN_T8 48u         // ... on an 8-lane group; above: a full warp

// `wl` = indices into elected[] of this group-size class (k_bin_elected).  Groups of fewer than 32
// lanes do not carry the 4096-bit function tables: a variable that reaches the function-table step
// with every neighbour mapped (varcore < 12; rare - only 12 variables are mapped per round) has
// no side effects yet (failed gate attempts restore their marks) and is handed to the 32-lane
// instance through `redo`.
template <int GS>
__global__ void __launch_bounds__(128) k_ve_phase1(GT<GS> g, const u32* __restrict__ wl, const u32* __restrict__ wlCount,
                                                   u32* __restrict__ redo, u32* redoCount) {
    constexpr int SLICE = GS == 32 ? VE_SLICE : VE_SLICE_SMALL;
    __shared__ u32 sh[128 / GS][SLICE];
    u32* out_c = sh[threadIdx.x / GS];
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 tid = wl[wi];
        const u32 x = g.elected[tid], p = V2L(x), n = p | 1u;
        const u32 np = g.otSize[p], nn = g.otSize[n];
        const u32* P = g.occurs + g.otStart[p];
        const u32* N = g.occurs + g.otStart[n];
        prefetchLists(g, P, np, N, nn);
        u32 pOrgs, nOrgs;
        if (g.k.in_mode) { u32 d; countOrgsLits(g, P, np, pOrgs, d); countOrgsLits(g, N, nn, nOrgs, d); }
        else { pOrgs = np; nOrgs = nn; }
        u32 elimType = 0, nElements = 0, nAddedCls = 0, nAddedLits = 0;
        bool eliminatedNow = false, record = false;
        if (!pOrgs || !nOrgs) { toblivionSave(g, p, n, pOrgs, nOrgs, P, np, N, nn); eliminatedNow = true; }
        else {
            const u32 def = findEquGate(g, p, n, P, np, N, nn);
            if (def) {
                // saveResolved(p, n, pOrgs, nOrgs, ...) elimination.cuh:552-594
                if (pOrgs > nOrgs) saveSide(g, N, nn, n, p, nOrgs); else saveSide(g, P, np, p, n, pOrgs);
                substituteSingle(g, p, n, def, P, np, N, nn);
                eliminatedNow = true;
            }
            else if ((pOrgs == 1 || nOrgs == 1) && !countPairs(g, 0, x, P, np, N, nn, 0, nElements, nAddedCls, nAddedLits)) {
                if (nAddedCls) { elimType = RES_MASK; record = true; }
                else { toblivionSave(g, p, n, pOrgs, nOrgs, P, np, N, nn); eliminatedNow = true; }
            }
            else {
                const u32 nClsBefore = pOrgs + nOrgs;
                elimType = 0; nElements = 0; nAddedCls = 0; nAddedLits = 0;
                if (nClsBefore > 2) {
                    if (nOrgs < g.k.sh_max_bve_out1 && findAOGate(g, n, N, nn, p, P, np, nClsBefore, out_c, nElements, nAddedCls, nAddedLits))
                        elimType = AOIX_MASK;
                    else if (!nAddedCls && pOrgs < g.k.sh_max_bve_out1 && findAOGate(g, p, P, np, n, N, nn, nClsBefore, out_c, nElements, nAddedCls, nAddedLits))
                        elimType = AOIX_MASK;
                }
                if (!elimType && nClsBefore > 3) {
                    // The ITE / XOR searches look for clauses that contain x or its negation, so they can walk x's OWN lists
                    // (in L1 after prefetchLists) instead of the shortest foreign list (cold: two L2 / DRAM round trips per
                    // probe) and find the same clause - the argument of k_ve_phase1_local.  Long lists keep the foreign walk.
                    const bool own = GS < 32 || np + nn <= 512u;
                    if (own) {
                        if (findITEGate<GS, true>(g, p, P, np, n, N, nn, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                        else if (!nAddedCls && findITEGate<GS, true>(g, n, N, nn, p, P, np, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                        else if (findXORGate<GS, true>(g, p, P, np, n, N, nn, nClsBefore, out_c, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                        else if (!nAddedCls && findXORGate<GS, true>(g, n, N, nn, p, P, np, nClsBefore, out_c, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                    } else if constexpr (GS == 32) {
                        if (findITEGate(g, p, P, np, n, N, nn, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                        else if (!nAddedCls && findITEGate(g, n, N, nn, p, P, np, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                        else if (findXORGate(g, p, P, np, n, N, nn, nClsBefore, out_c, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                        else if (!nAddedCls && findXORGate(g, n, N, nn, p, P, np, nClsBefore, out_c, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                    }
                }
                bool funHit = false;
                if (g.k.ve_fun_en && !elimType && nClsBefore > 2) {
                    if constexpr (GS == 32) funHit = findFunGate(g, p, n, nClsBefore, P, np, N, nn, nElements, nAddedCls, nAddedLits);
                    else if (funPossible(g, p, P, np) && funPossible(g, n, N, nn)) {
                        if (LANE == 0) redo[atomicAdd(redoCount, 1u)] = tid;
                        GSYNC();
                        continue;
                    }
                }
                if (funHit) elimType = CORE_MASK;
                else if (!elimType && !nAddedCls && !countPairs(g, 1, x, P, np, N, nn, nClsBefore, nElements, nAddedCls, nAddedLits))
                    elimType = RES_MASK;
                if (!nAddedCls) { toblivionSave(g, p, n, pOrgs, nOrgs, P, np, N, nn); eliminatedNow = true; }
                else if (elimType) record = true;
            }
        }
        if (LANE == 0) {
            if (record) {
                g.veType[tid] = ENCODEVARINFO(elimType, nAddedCls, nAddedLits);
                g.veUcnt[tid] = nElements; g.veRpos[tid] = nAddedCls; g.veRref[tid] = (u64)nAddedLits + (u64)NBUCKETS * nAddedCls;
            } else { g.veType[tid] = 0; g.veUcnt[tid] = 0; g.veRpos[tid] = 0; g.veRref[tid] = 0; }
            if (eliminatedNow) g.eliminated[x] |= MELTING_MASK;
        }
        GSYNC();
    }
}

// ------------------------------------------------------------------ BVE phase 1 on a local copy (4- and 8-lane classes)
// The decision tree of bounded.cuh:282-394 walks the same few clauses again and again: every gate search re-reads
// headers and literals through list entry -> header -> literals, three dependent loads per step, and the ITE / XOR
// searches walk FOREIGN lists on top.  A variable of these classes has at most BIN_T8 clauses: they are gathered ONCE
// (all lanes in parallel, the three loads of different clauses in flight together) into a shared-memory copy, the whole
// tree - marks included - runs on the copy, and only the outcome goes back: changed molten marks (phase 3 reads them),
// then the terminal action (witness + deletion, equivalence substitution) on the global clauses.
// A variable with a clause longer than VE_LOCAL_K literals is handed to the 32-lane kernel (`redo`), like the
// function-table candidates.
#define VE_LOCAL_K 8
template <int GS>
__global__ void __launch_bounds__(128) k_ve_phase1_local(GT<GS> g, const u32* __restrict__ wl, const u32* __restrict__ wlCount,
                                                         u32* __restrict__ redo, u32* redoCount) {
    constexpr int MAXC = GS == 4 ? (int)BIN_T4 : (int)BIN_T8;
    constexpr int NG = 128 / GS;
    constexpr int CNT = MAXC / GS;
    __shared__ uint4 s_h[NG][MAXC];
    __shared__ u32 s_l[NG][MAXC * VE_LOCAL_K];
    __shared__ u32 s_w0[NG][MAXC];
    __shared__ u32 s_out[NG][VE_SLICE_SMALL];
    __shared__ u32 s_iota[MAXC];
    if (threadIdx.x < MAXC) s_iota[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const u32 gidx = threadIdx.x / GS;
    uint4* lh = s_h[gidx]; u32* ll = s_l[gidx]; u32* w0 = s_w0[gidx]; u32* out_c = s_out[gidx];
    GT<GS> lg = g;                 // the same engine state, clause store = the local copy, clause "index" = slot
    lg.hdr = lh; lg.pool = ll;
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 tid = wl[wi];
        const u32 x = g.elected[tid], p = V2L(x), n = p | 1u;
        const u32 np = g.otSize[p], nn = g.otSize[n], tot = np + nn;
        const u32* Pg = g.occurs + g.otStart[p];
        const u32* Ng = g.occurs + g.otStart[n];
        // ---- gather: entries, then headers, then literals, each as one batch of independent loads
        u32 ci[CNT]; uint4 h[CNT];
        bool big = tot > (u32)MAXC;
#pragma unroll
        for (int q = 0; q < CNT; q++) { const u32 j = LANE + q * GS; ci[q] = j < tot && !big ? (j < np ? Pg[j] : Ng[j - np]) : NOVAR; }
#pragma unroll
        for (int q = 0; q < CNT; q++) { h[q] = ci[q] != NOVAR ? g.hdr[ci[q]] : make_uint4(0, 0, 0, 0); big |= h[q].y > VE_LOCAL_K; }
        if (__any_sync(FULL, big)) {
            if (LANE == 0) redo[atomicAdd(redoCount, 1u)] = tid;
            GSYNC();
            continue;
        }
        GSYNC();   // the previous variable's readers of this slice are done
#pragma unroll
        for (int q = 0; q < CNT; q++) {
            const u32 j = LANE + q * GS;
            if (ci[q] != NOVAR) {
                const u32* src = g.pool + h[q].x;
                u32 lv[VE_LOCAL_K];
#pragma unroll
                for (int k = 0; k < VE_LOCAL_K; k++) lv[k] = (u32)k < h[q].y ? src[k] : 0u;
#pragma unroll
                for (int k = 0; k < VE_LOCAL_K; k++) ll[j * VE_LOCAL_K + k] = lv[k];
                lh[j] = make_uint4(j * VE_LOCAL_K, h[q].y, h[q].z, h[q].w);
                w0[j] = h[q].w;
            }
        }
        GSYNC();
        const u32* P = s_iota; const u32* N = s_iota + np;
        // ---- the decision tree on the copy (same order and guards as k_ve_phase1)
        u32 pOrgs, nOrgs;
        if (g.k.in_mode) { u32 d; countOrgsLits(lg, P, np, pOrgs, d); countOrgsLits(lg, N, nn, nOrgs, d); }
        else { pOrgs = np; nOrgs = nn; }
        u32 elimType = 0, nElements = 0, nAddedCls = 0, nAddedLits = 0, def = 0;
        bool oblivion = false, record = false, handOver = false;
        if (!pOrgs || !nOrgs) oblivion = true;
        else {
            def = findEquGate(lg, p, n, P, np, N, nn);
            if (def) {}
            else if ((pOrgs == 1 || nOrgs == 1) && !countPairs(lg, 0, x, P, np, N, nn, 0, nElements, nAddedCls, nAddedLits)) {
                if (nAddedCls) { elimType = RES_MASK; record = true; }
                else oblivion = true;
            }
            else {
                const u32 nClsBefore = pOrgs + nOrgs;
                elimType = 0; nElements = 0; nAddedCls = 0; nAddedLits = 0;
                if (nClsBefore > 2) {
                    if (nOrgs < g.k.sh_max_bve_out1 && findAOGate(lg, n, N, nn, p, P, np, nClsBefore, out_c, nElements, nAddedCls, nAddedLits))
                        elimType = AOIX_MASK;
                    else if (!nAddedCls && pOrgs < g.k.sh_max_bve_out1 && findAOGate(lg, p, P, np, n, N, nn, nClsBefore, out_c, nElements, nAddedCls, nAddedLits))
                        elimType = AOIX_MASK;
                }
                if (!elimType && nClsBefore > 3) {
                    if (findITEGate<GS, true>(lg, p, P, np, n, N, nn, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                    else if (!nAddedCls && findITEGate<GS, true>(lg, n, N, nn, p, P, np, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                    else if (findXORGate<GS, true>(lg, p, P, np, n, N, nn, nClsBefore, out_c, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                    else if (!nAddedCls && findXORGate<GS, true>(lg, n, N, nn, p, P, np, nClsBefore, out_c, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
                }
                if (g.k.ve_fun_en && !elimType && nClsBefore > 2 && funPossible(lg, p, P, np) && funPossible(lg, n, N, nn)) handOver = true;
                if (!handOver) {
                    if (!elimType && !nAddedCls && !countPairs(lg, 1, x, P, np, N, nn, nClsBefore, nElements, nAddedCls, nAddedLits))
                        elimType = RES_MASK;
                    if (!nAddedCls) oblivion = true;
                    else if (elimType) record = true;
                }
            }
        }
        if (handOver) {   // no side effect outside the copy yet: failed gate attempts restored their marks
            if (LANE == 0) redo[atomicAdd(redoCount, 1u)] = tid;
            GSYNC();
            continue;
        }
        // ---- outcome: marks that changed, then the terminal action on the global clauses
        GSYNC();
#pragma unroll
        for (int q = 0; q < CNT; q++) {
            const u32 j = LANE + q * GS;
            if (ci[q] != NOVAR) { const u32 w = lh[j].w; if (w != w0[j]) g.hdr[ci[q]].w = w; }
        }
        GSYNC();
        bool eliminatedNow = false;
        if (oblivion) { toblivionSave(g, p, n, pOrgs, nOrgs, Pg, np, Ng, nn); eliminatedNow = true; }
        else if (def) {
            if (pOrgs > nOrgs) saveSide(g, Ng, nn, n, p, nOrgs); else saveSide(g, Pg, np, p, n, pOrgs);
            substituteSingle(g, p, n, def, Pg, np, Ng, nn);
            eliminatedNow = true;
        }
        if (LANE == 0) {
            if (record) {
                g.veType[tid] = ENCODEVARINFO(elimType, nAddedCls, nAddedLits);
                g.veUcnt[tid] = nElements; g.veRpos[tid] = nAddedCls; g.veRref[tid] = (u64)nAddedLits + (u64)NBUCKETS * nAddedCls;
            } else { g.veType[tid] = 0; g.veUcnt[tid] = 0; g.veRpos[tid] = 0; g.veRref[tid] = 0; }
            if (eliminatedNow) g.eliminated[x] |= MELTING_MASK;
        }
        GSYNC();
    }
}

// ------------------------------------------------------------------ BVE phase 3 (bounded.cuh:488-544, 171-280)
// start values of the scans: CNF sizes before this BVE (elimination.cu:82-92)
struct VEBase { u32 numCls0, poolUsed0; u64 dataSize0; };

template <int GS>
__global__ void __launch_bounds__(128) k_ve_phase3(GT<GS> g, VEBase vb, u32* __restrict__ survivorFlag, const u32* __restrict__ wl,
                                                   const u32* __restrict__ wlCount) {
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 tid = wl[wi];
        const u32 x = g.elected[tid];
        const u32 xinfo = g.veType[tid], elimType = RECOVERTYPE(xinfo);
        if (elimType) {
            const u32 p = V2L(x), n = p | 1u;
            const u32 nAddedCls = RECOVERADDEDCLS(xinfo), nAddedLits = RECOVERADDEDLITS(xinfo);
            const u32 addedPos = g.veRpos[tid];
            const u64 addedRef = g.veRref[tid];
            const u32 np = g.otSize[p], nn = g.otSize[n];
            const u32* P = g.occurs + g.otStart[p];
            const u32* N = g.occurs + g.otStart[n];
            bool safe = ((u64)addedPos + nAddedCls <= g.k.refsCap) &&
                        (addedRef + nAddedLits + (u64)NBUCKETS * nAddedCls <= g.k.dataCap);
            if (safe) {   // the logical capacities can pass the arena after a GC (api.cu: prepareLoad): fail loudly, never write outside
                const u64 poolEnd = (u64)vb.poolUsed0 + ((addedRef - vb.dataSize0) - (u64)NBUCKETS * (addedPos - vb.numCls0)) + nAddedLits;
                if ((u64)addedPos + nAddedCls > g.k.physC || poolEnd > g.k.physW) { safe = false; if (LANE == 0) atomicOr(&g.dc->flags, 64u); }
            }
            if (safe) {
                // saveResolved(p, n, cnf, poss, negs) elimination.cuh:505-550 : side chosen by list sizes
                u32 c1, l1;
                if (np > nn) { countOrgsLits(g, N, nn, c1, l1); saveSide(g, N, nn, n, p, c1); }
                else { countOrgsLits(g, P, np, c1, l1); saveSide(g, P, np, p, n, c1); }
                // emission: pairs in pos-major / neg-minor order, 32 at a time
                u32 clsOut = addedPos;
                u32 poolOut = vb.poolUsed0 + (u32)((addedRef - vb.dataSize0) - (u64)NBUCKETS * (addedPos - vb.numCls0));
                const u32 checksum = addedPos + nAddedCls;
                const u32 nUnitsExp = g.veUcnt[tid];
                u32 ucursor = nUnitsExp ? reserveUnits(g, nUnitsExp) : 0;
                const u64 total = (u64)np * nn;
                for (u64 t0 = 0; t0 < total && clsOut < checksum; t0 += GS) {
                    const u64 t = t0 + LANE;
                    int rsize = 0;
                    uint4 hi = make_uint4(0, 0, 0, 0), hj = hi;
                    if (t < total) {
                        const u32 i = (u32)(t / nn), j = (u32)(t - (u64)i * nn);
                        hi = g.hdr[P[i]];
                        if (!C_LEARNT(hi.w)) {
                            hj = g.hdr[N[j]];
                            bool take = !C_LEARNT(hj.w);
                            if (elimType == AOIX_MASK) take = take && ((C_MOLTEN(hi.w) != 0) != (C_MOLTEN(hj.w) != 0));
                            else if (elimType == CORE_MASK) take = take && !(C_MOLTEN(hi.w) && C_MOLTEN(hj.w));
                            if (take) rsize = mergeLen(g.pool + hi.x, (int)hi.y, g.pool + hj.x, (int)hj.y, x);
                        }
                    }
                    const bool isCls = rsize > 1, isUnit = rsize == 1;
                    const u32 mc = BALLOT(isCls), mu = BALLOT(isUnit);
                    const u32 myCls = __popc(mc & LTMASK);
                    const u32 wordsIncl = gIncl<GS>(isCls ? (u32)rsize : 0u);
                    const u32 myWords = wordsIncl - (isCls ? (u32)rsize : 0u);
                    const u32 totWords = __shfl_sync(FULL, wordsIncl, GS - 1, GS);
                    if (isCls && clsOut + myCls < checksum) {
                        u32 sig;
                        u32* out = g.pool + poolOut + myWords;
                        mergeOut(g.pool + hi.x, (int)hi.y, g.pool + hj.x, (int)hj.y, x, out, sig);
                        // new SCLAUSE: ORIGINAL, added (bounded.cuh:86-120)
                        g.hdr[clsOut + myCls] = make_uint4(poolOut + myWords, (u32)rsize, sig, CB_ADDED);
                    }
                    if (isUnit) {
                        u32 sig, lit;
                        mergeOut(g.pool + hi.x, (int)hi.y, g.pool + hj.x, (int)hj.y, x, &lit, sig);
                        const u32 slot = ucursor + __popc(mu & LTMASK);
                        if (slot < g.unitsCap) g.units[slot] = lit; else atomicOr(&g.dc->flags, 2u);
                    }
                    clsOut += __popc(mc); poolOut += totWords; ucursor += __popc(mu);
                }
                GSYNC();
                deleteAll(g, P, np);
                deleteAll(g, N, nn);
                if (LANE == 0) {
                    g.otSize[p] = 0; g.otSize[n] = 0;
                    g.eliminated[x] |= (MELTING_MASK | ADDING_MASK);
                    atomicMax(&g.dc->lastElimID, (int)tid);
                }
                GSYNC();
            }
            else {
                if (LANE == 0) atomicOr(&g.dc->flags, 4u);
                if (elimType != RES_MASK) freezeClauses(g, P, np, N, nn);
            }
        }
        if (LANE == 0) survivorFlag[tid] = g.eliminated[x] ? 0u : 1u;
        GSYNC();
    }
}

// resizeCNF_k (cnf.cu:55-79)
__global__ void k_ve_resize(G g, VEBase vb) {
    const int last = g.dc->lastElimID;
    if (last >= 0) {
        const u32 info = g.veType[last];
        const u32 cl = RECOVERADDEDCLS(info), li = RECOVERADDEDLITS(info);
        const u32 pos = g.veRpos[last];
        const u64 ref = g.veRref[last];
        g.dc->numCls = pos + cl;
        g.dc->dataSize = ref + li + (u64)NBUCKETS * cl;
        g.dc->poolUsed = vb.poolUsed0 + (u32)((ref - vb.dataSize0) - (u64)NBUCKETS * (pos - vb.numCls0)) + li;
        g.dc->addedCls = pos + cl - vb.numCls0;
    } else g.dc->addedCls = 0;
}
__global__ void k_ve_reset(DevCounters* dc) { dc->lastElimID = -1; dc->addedCls = 0; }
__global__ void k_select_scatter(const u32* __restrict__ src, const u32* __restrict__ flag, const u32* __restrict__ pos, u32 n,
                                 u32* __restrict__ dst) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) dst[pos[i]] = src[i];
}
__global__ void k_copy_u32(const u32* __restrict__ src, u32* __restrict__ dst, const u32* n) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < *n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------ SUB (subsume.cuh)
// selfsub signature test (elimination.cuh:67-71)
__device__ __forceinline__ bool selfsubSig(u32 A, u32 B) {
    const u32 Bt = B | ((B & 0xAAAAAAAAu) >> 1) | ((B & 0x55555555u) << 1);
    return !(A & ~Bt);
}
// subsume.cuh:132-176
__device__ __forceinline__ bool selfsubMerge(const u32* d1, int n1, const u32* d2, int n2, u32 x, u32 fx) {
    int i1 = 0, i2 = 0, sub = 0; bool self = false;
    while (i1 < n1 && i2 < n2) {
        const u32 lit1 = d1[i1], lit2 = d2[i2];
        if (lit1 == fx) i1++;
        else if (lit2 == x) { self = true; i2++; }
        else if (lit1 < lit2) i1++;
        else if (lit2 < lit1) i2++;
        else { sub++; i1++; i2++; }
    }
    if (sub + 1 == n1) {
        if (self) return true;
        while (i2 < n2) { if (d2[i2] == x) return true; i2++; }
    }
    return false;
}
// subsume.cuh:50-75
__device__ __forceinline__ bool subMerge(const u32* d1, int n1, const u32* d2, int n2) {
    int i1 = 0, i2 = 0, sub = 0;
    while (i1 < n1 && i2 < n2) {
        const u32 lit1 = d1[i1], lit2 = d2[i2];
        if (lit1 < lit2) i1++;
        else if (lit2 < lit1) i2++;
        else { sub++; i1++; i2++; }
    }
    return sub == n1;
}

// one side of sub_k: every clause of M against the other list O, then against its predecessors in M
template <int GS> __device__ u32 subSide(GT<GS>& g, const u32* M, u32 nm, const u32* O, u32 no, u32 x, u32 fx) {
    u32 nUnits = 0;
    for (u32 i = 0; i < nm; i++) {
        const u32 ci = M[i];
        uint4 h = g.hdr[ci];
        if ((int)h.y > SUB_MAX_CL_SIZE) break;
        if (C_DELETED(h.w)) continue;
        // selfsubsume (subsume.cuh:344-400): first neg clause, in list order, that strengthens cand
        bool hit = false;
        for (u32 base = 0; base < no; base += GS) {
            const u32 j = base + LANE;
            bool brk = false, ok = false;
            if (j < no) {
                const uint4 hj = g.hdr[O[j]];
                if (hj.y > h.y) brk = true;
                else if (!C_DELETED(hj.w) && !C_MOLTEN(hj.w) && hj.y > 1 && selfsubSig(hj.z, h.z) &&
                         selfsubMerge(g.pool + hj.x, (int)hj.y, g.pool + h.x, (int)h.y, x, fx)) ok = true;
            }
            const u32 bm = BALLOT(brk), om = BALLOT(ok);
            const u32 fb = bm ? (u32)__ffs(bm) - 1 : 32u, fo = om ? (u32)__ffs(om) - 1 : 32u;
            if (fo < fb) { hit = true; break; }
            if (bm) break;
        }
        if (hit) {  // strengthen (subsume.cuh:269-291) + melt
            if (LANE == 0) {
                u32* l = g.pool + h.x;
                u32 n = 0;
                for (u32 k = 0; k < h.y; k++) if (l[k] != x) l[n++] = l[k];
                h.y = h.y - 1;
                if (h.y > 1) {
                    h.z = sigOf(l, (int)h.y);
                    if (C_LEARNT(h.w)) {  // bumpShrunken (subsume.cuh:226-238)
                        const int old_lbd = (int)(h.w >> CB_LBD_SHIFT);
                        if (old_lbd > LBD_TIER1) {
                            const int new_lbd = min((int)h.y - 1, old_lbd);
                            if (new_lbd < old_lbd) h.w = (h.w & ((1u << CB_LBD_SHIFT) - 1) & ~CB_USAGE_MASK) | ((u32)new_lbd << CB_LBD_SHIFT) | (USAGET3 << CB_USAGE_SHIFT);
                        }
                    }
                }
                h.w |= CB_MOLTEN;
                g.hdr[ci] = h;
            }
            GSYNC();
            h = g.hdr[ci];
            if (h.y == 1) nUnits++;
        }
        // subsume (subsume.cuh:305-342): first earlier clause of the same list that subsumes cand
        const bool candMolten = C_MOLTEN(h.w) != 0;
        for (u32 base = 0; base < i; base += GS) {
            const u32 j = base + LANE;
            bool ok = false;
            if (j < i) {
                const uint4 hj = g.hdr[M[j]];
                if (!C_DELETED(hj.w) && !(candMolten && hj.y > h.y) && hj.y > 1 && SUBSIG(hj.z, h.z) &&
                    subMerge(g.pool + hj.x, (int)hj.y, g.pool + h.x, (int)h.y)) ok = true;
            }
            const u32 om = BALLOT(ok);
            if (om) {
                const u32 cj = M[base + __ffs(om) - 1];
                if (LANE == 0) {
                    const u32 wj = g.hdr[cj].w;
                    if (C_LEARNT(wj) && C_ORIGINAL(h.w)) g.hdr[cj].w = wj & ~CB_ST_MASK;
                    g.hdr[ci].w = (h.w & ~CB_ST_MASK) | CB_DELETED;
                }
                GSYNC();
                break;
            }
        }
    }
    return nUnits;
}
// updateOL (subsume.cuh:293-303): drop molten (un-melting them) and deleted clauses, keep order
template <int GS> __device__ void updateOL(GT<GS>& g, u32 lit) {
    const u32 n = g.otSize[lit];
    if (!n) return;
    u32* list = g.occurs + g.otStart[lit];
    u32 out = 0;
    for (u32 base = 0; base < n; base += GS) {
        const u32 j = base + LANE;
        u32 ci = 0; bool keep = false;
        if (j < n) {
            ci = list[j];
            const u32 w = g.hdr[ci].w;
            if (C_MOLTEN(w)) { if (!g.k.proof_en) g.hdr[ci].w = w & ~CB_MOLTEN; }   // proof mode: k_proof_stream prints, then freezes
            else if (!C_DELETED(w)) keep = true;
        }
        const u32 m = BALLOT(keep);
        GSYNC();
        if (keep) list[out + __popc(m & LTMASK)] = ci;
        out += __popc(m);
        GSYNC();
    }
    if (LANE == 0) g.otSize[lit] = out;
    GSYNC();
}

template <int GS>
__global__ void __launch_bounds__(128) k_sub(GT<GS> g, const u32* __restrict__ wl, const u32* __restrict__ wlCount) {
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 x = g.elected[wl[wi]], p = V2L(x), n = p | 1u;
        const u32 np = g.otSize[p], nn = g.otSize[n];
        if (np > g.k.sub_max_occurs || nn > g.k.sub_max_occurs) continue;
        const u32* P = g.occurs + g.otStart[p];
        const u32* N = g.occurs + g.otStart[n];
        prefetchLists(g, P, np, N, nn);
        const u32 nPosUnits = subSide(g, P, np, N, nn, p, n);
        const u32 nNegUnits = subSide(g, N, nn, P, np, n, p);
        if (nPosUnits || nNegUnits) {
            u32 cursor = reserveUnits(g, nPosUnits + nNegUnits);
            if (nPosUnits) appendUnits(g, P, np, cursor);
            if (nNegUnits) appendUnits(g, N, nn, cursor);
        }
        updateOL(g, p);
        updateOL(g, n);
    }
}

// SUB on a local copy (4- and 8-lane classes): same idea as k_ve_phase1_local - gather the variable's clauses once, run
// strengthening and subsumption of both sides on the shared-memory copy, write back only the clauses that changed
// (header; literals when one was removed), then units and updateOL on the global lists.  A variable with a clause longer
// than VE_LOCAL_K literals goes to `redo`, processed by k_sub<32>.
template <int GS>
__global__ void __launch_bounds__(128) k_sub_local(GT<GS> g, const u32* __restrict__ wl, const u32* __restrict__ wlCount,
                                                   u32* __restrict__ redo, u32* redoCount) {
    constexpr int MAXC = GS == 4 ? (int)BIN_T4 : (int)BIN_T8;
    constexpr int NG = 128 / GS;
    constexpr int CNT = MAXC / GS;
    __shared__ uint4 s_h[NG][MAXC];
    __shared__ u32 s_l[NG][MAXC * VE_LOCAL_K];
    __shared__ u32 s_iota[MAXC];
    if (threadIdx.x < MAXC) s_iota[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const u32 gidx = threadIdx.x / GS;
    uint4* lh = s_h[gidx]; u32* ll = s_l[gidx];
    GT<GS> lg = g;
    lg.hdr = lh; lg.pool = ll;
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 tid = wl[wi];
        const u32 x = g.elected[tid], p = V2L(x), n = p | 1u;
        const u32 np = g.otSize[p], nn = g.otSize[n], tot = np + nn;
        if (np > g.k.sub_max_occurs || nn > g.k.sub_max_occurs) continue;
        const u32* Pg = g.occurs + g.otStart[p];
        const u32* Ng = g.occurs + g.otStart[n];
        u32 ci[CNT]; uint4 h[CNT];
        bool big = tot > (u32)MAXC;
#pragma unroll
        for (int q = 0; q < CNT; q++) { const u32 j = LANE + q * GS; ci[q] = j < tot && !big ? (j < np ? Pg[j] : Ng[j - np]) : NOVAR; }
#pragma unroll
        for (int q = 0; q < CNT; q++) { h[q] = ci[q] != NOVAR ? g.hdr[ci[q]] : make_uint4(0, 0, 0, 0); big |= h[q].y > VE_LOCAL_K; }
        if (__any_sync(FULL, big)) {
            if (LANE == 0) redo[atomicAdd(redoCount, 1u)] = tid;
            GSYNC();
            continue;
        }
        GSYNC();
#pragma unroll
        for (int q = 0; q < CNT; q++) {
            const u32 j = LANE + q * GS;
            if (ci[q] != NOVAR) {
                const u32* src = g.pool + h[q].x;
                u32 lv[VE_LOCAL_K];
#pragma unroll
                for (int k = 0; k < VE_LOCAL_K; k++) lv[k] = (u32)k < h[q].y ? src[k] : 0u;
#pragma unroll
                for (int k = 0; k < VE_LOCAL_K; k++) ll[j * VE_LOCAL_K + k] = lv[k];
                lh[j] = make_uint4(j * VE_LOCAL_K, h[q].y, h[q].z, h[q].w);
            }
        }
        GSYNC();
        const u32* P = s_iota; const u32* N = s_iota + np;
        const u32 nPosUnits = subSide(lg, P, np, N, nn, p, n);
        const u32 nNegUnits = subSide(lg, N, nn, P, np, n, p);
        GSYNC();
        // write back what changed: a strengthened clause lost one literal (new size, signature, marks), a subsumed one is deleted
#pragma unroll
        for (int q = 0; q < CNT; q++) {
            const u32 j = LANE + q * GS;
            if (ci[q] != NOVAR) {
                const uint4 nh = lh[j];
                if (nh.y != h[q].y) { u32* dst = g.pool + h[q].x; for (u32 k = 0; k < nh.y; k++) dst[k] = ll[j * VE_LOCAL_K + k]; }
                if (nh.y != h[q].y || nh.z != h[q].z || nh.w != h[q].w) g.hdr[ci[q]] = make_uint4(h[q].x, nh.y, nh.z, nh.w);
            }
        }
        GSYNC();
        if (nPosUnits || nNegUnits) {
            u32 cursor = reserveUnits(g, nPosUnits + nNegUnits);
            if (nPosUnits) appendUnits(g, Pg, np, cursor);
            if (nNegUnits) appendUnits(g, Ng, nn, cursor);
        }
        updateOL(g, p);
        updateOL(g, n);
    }
}
// the variables handed over by k_sub_local
// ------------------------------------------------------------------ BCE (blocked.cuh:26-97)
template <int GS>
__global__ void __launch_bounds__(128) k_bce(GT<GS> g, const u32* __restrict__ wl, const u32* __restrict__ wlCount) {
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 x = g.elected[wl[wi]], p = V2L(x), n = p | 1u;
        const u32 np = g.otSize[p], nn = g.otSize[n];
        if (np > g.k.bce_max_occurs || nn > g.k.bce_max_occurs) continue;
        const u32* P = g.occurs + g.otStart[p];
        const u32* N = g.occurs + g.otStart[n];
        for (u32 i = 0; i < nn; i++) {
            const u32 ci = N[i]; const uint4 hi = g.hdr[ci];
            if (C_DELETED(hi.w) || C_LEARNT(hi.w)) continue;
            bool nonTaut = false;
            for (u32 base = 0; base < np && !nonTaut; base += GS) {
                const u32 j = base + LANE;
                bool nt = false;
                if (j < np) {
                    const uint4 hj = g.hdr[P[j]];
                    if (!C_DELETED(hj.w) && !C_LEARNT(hj.w))
                        nt = !isTautology(g.pool + hj.x, (int)hj.y, g.pool + hi.x, (int)hi.y, x);
                }
                nonTaut = __any_sync(FULL, nt);
            }
            if (!nonTaut) {
                u32* saved = jumpResolved(g, hi.y + 1);
                if (saved && LANE == 0) saveClause(g, saved, hi, n);
                markDeleted(g, ci);
            }
        }
    }
}

// ------------------------------------------------------------------ ERE, lane per resolvent
// Same semantics as ere_k (redundancy.cuh:99-174), different work mapping:
//   * one warp per elected variable, one LANE per (pos, neg) pair - the 32 lanes build 32
//     different resolvents at once instead of rebuilding the same one;
//   * the resolvent is never stored: one merge pass yields its length, first/last literal,
//     signature and the literal with the shortest occurrence list;
//   * a clause equal to the resolvent has the same (size, first, last, sig), which is exactly the
//     key the occurrence lists are sorted by (key.cuh:67-83), so the candidates are found by a
//     binary search over the sorted list instead of scanning it: ~log2(n) key probes of 16 B
//     replace n header reads;
//   * the reference's rule "lane t checks entries t, t+32, ... and deletes its first match" is
//     kept exactly: within the run of key-equal entries, the first matching entry of every
//     residue class (position mod 32) is deleted.
// Algorithmic bytes per resolvent: 2 clause reads (L1/L2 resident per variable) + 4 B per
// resolvent literal (list sizes) + 20 B per binary-search probe.
//   * before any list is touched the resolvent's key is looked up in a Bloom filter over the keys of
//     all live clauses (one bit each, <= 16 bits per clause: L2 resident).  A clause equal to the
//     resolvent has the same key, so a clear bit proves there is none; only the few hits (false
//     positives included) go on to the binary search.  The filter never changes a result.
__device__ __forceinline__ u32 keyHash(u32 sz, u32 first, u32 last, u32 sig) {
    u32 h = sz * 0x9E3779B1u;
    h ^= first * 0x85EBCA6Bu + 0x7F4A7C15u + (h << 6) + (h >> 2);
    h ^= last * 0xC2B2AE35u + (h << 6) + (h >> 2);
    h ^= sig * 0x27D4EB2Fu + (h << 6) + (h >> 2);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
// Streams key[] only: the counting pass of this round (k_ot_count) wrote a zero key for every deleted slot.
__global__ void __launch_bounds__(256) k_ere_bloom(const uint4* __restrict__ key, u32 n, u32* __restrict__ bloom, u32 mask) {
    __shared__ u32 sizes[8];   // 256-bit mask of the clause sizes seen by this CTA; stored after the filter words
    if (threadIdx.x < 8) sizes[threadIdx.x] = 0;
    __syncthreads();
    u32 seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 k = key[i];
        if (!k.x) continue;
        const u32 h = keyHash(k.x, k.y, k.z, k.w) & mask;
        atomicOr(&bloom[h >> 5], 1u << (h & 31u));
        const u32 sb = k.x < 255u ? k.x : 255u;
#pragma unroll
        for (int w = 0; w < 8; w++) if ((sb >> 5) == (u32)w) seen[w] |= 1u << (sb & 31u);
    }
#pragma unroll
    for (int w = 0; w < 8; w++) {
        u32 m = seen[w];
#pragma unroll
        for (int o = 16; o; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
        if ((threadIdx.x & 31u) == 0 && m) atomicOr(&sizes[w], m);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sizes[threadIdx.x]) atomicOr(&bloom[(mask >> 5) + 1 + threadIdx.x], sizes[threadIdx.x]);
}
__device__ __forceinline__ int keyCmp(const uint4 k, u32 sz, u32 first, u32 last, u32 sig) {
    if (k.x != sz) return k.x < sz ? -1 : 1;
    if (k.y != first) return k.y < first ? -1 : 1;
    if (k.z != last) return k.z < last ? -1 : 1;
    if (k.w != sig) return k.w < sig ? -1 : 1;
    return 0;
}
// does the resolvent of a and b on x equal cand[0..len)?  (same merge as mergeOut, streamed)
__device__ __forceinline__ bool resolventEquals(const u32* a, int n1, const u32* b, int n2, u32 x, const u32* cand) {
    int it1 = 0, it2 = 0, k = 0;
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = LABS(lit1), v2 = LABS(lit2);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (v1 < v2) { it1++; if (cand[k++] != lit1) return false; }
        else if (v2 < v1) { it2++; if (cand[k++] != lit2) return false; }
        else { it1++; it2++; if (cand[k++] != lit1) return false; }
    }
    while (it1 < n1) { const u32 l = a[it1++]; if (LABS(l) != x && cand[k++] != l) return false; }
    while (it2 < n2) { const u32 l = b[it2++]; if (LABS(l) != x && cand[k++] != l) return false; }
    return true;
}

// One merge pass over the resolvent of a and b on variable v: length, first / last literal and
// signature - everything the filters and the key search need.  Returns false for a tautology.
__device__ __forceinline__ bool ereKey(const u32* a, int n1, const u32* b, int n2, u32 v, u32& len, u32& first, u32& last, u32& sig) {
    int it1 = 0, it2 = 0;
    len = 0; first = 0; last = 0; sig = 0;
#define ERE_EMIT(L_) do { const u32 l_ = (L_); if (!len) first = l_; last = l_; len++; sig |= MAPHASH(l_); } while (0)
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = LABS(lit1), v2 = LABS(lit2);
        if (v1 == v) it1++;
        else if (v2 == v) it2++;
        else if (IS_TAUT(lit1, lit2)) return false;
        else if (v1 < v2) { it1++; ERE_EMIT(lit1); }
        else if (v2 < v1) { it2++; ERE_EMIT(lit2); }
        else { it1++; it2++; ERE_EMIT(lit1); }
    }
    while (it1 < n1) { const u32 l = a[it1++]; if (LABS(l) != v) ERE_EMIT(l); }
    while (it2 < n2) { const u32 l = b[it2++]; if (LABS(l) != v) ERE_EMIT(l); }
#undef ERE_EMIT
    return true;
}
// the literal of the (non-tautological) resolvent with the shortest occurrence list (forward_equ, redundancy.cuh:99-112)
__device__ __forceinline__ u32 ereBest(const u32* __restrict__ otSize, const u32* a, int n1, const u32* b, int n2, u32 v, u32& minsize) {
    int it1 = 0, it2 = 0;
    u32 best = 0;
    minsize = 0xFFFFFFFFu;
#define ERE_EMIT(L_) do { const u32 l_ = (L_); const u32 s_ = otSize[l_]; if (s_ < minsize) { minsize = s_; best = l_; } } while (0)
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = LABS(lit1), v2 = LABS(lit2);
        if (v1 == v) it1++;
        else if (v2 == v) it2++;
        else if (v1 < v2) { it1++; ERE_EMIT(lit1); }
        else if (v2 < v1) { it2++; ERE_EMIT(lit2); }
        else { it1++; it2++; ERE_EMIT(lit1); }
    }
    while (it1 < n1) { const u32 l = a[it1++]; if (LABS(l) != v) ERE_EMIT(l); }
    while (it2 < n2) { const u32 l = b[it2++]; if (LABS(l) != v) ERE_EMIT(l); }
#undef ERE_EMIT
    return best;
}
// key search in the SORTED list of `best` + the reference's residue-class deletion rule
template <int GS>
__device__ __forceinline__ void ereSearchDelete(GT<GS>& g, const u32* a, int n1, const u32* b, int n2, u32 v, u32 len, u32 first, u32 last,
                                                u32 sig, u32 type, u32 best, u32 minsize) {
    const u32* list = g.occurs + g.otStart[best];
    u32 lo = 0, hi = minsize;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (keyCmp(g.key[list[mid]], len, first, last, sig) < 0) lo = mid + 1; else hi = mid;
    }
    u32 done = 0;
    for (u32 e = lo; e < minsize; e++) {
        const u32 ci = list[e];
        if (keyCmp(g.key[ci], len, first, last, sig) != 0) break;
        const u32 r = e & 31u;
        if ((done >> r) & 1u) continue;
        const uint4 h = g.hdr[ci];
        if ((C_LEARNT(h.w) || (h.w & CB_ST_MASK) == type) && !C_DELETED(h.w) && h.y == len &&
            resolventEquals(a, n1, b, n2, v, g.pool + h.x)) {
            g.hdr[ci].w = (h.w & ~CB_ST_MASK) | CB_DELETED;
            done |= 1u << r;
        }
    }
}

__device__ __forceinline__ void ereAdvance(u32& i, u32& j, u32 step, u32 fs) {
    j += step;
    if (fs >= step) { if (j >= fs) { j -= fs; i++; } }   // at most one wrap
    else { i += j / fs; j %= fs; }
}
// Phase A: enumerate the resolvents, keep the few that pass the filters.  With `queue` the survivors
// are only recorded (clause pair + variable) and the list they will be searched in is flagged, so
// that ONLY those lists have to be sorted before phase B (k_ere_apply); without it the search runs
// here, on lists the caller has sorted already.
struct EreQueue { u32* items; u32 cap; u32* count; unsigned char* need; u32* overflow; };
template <int GS>
__global__ void __launch_bounds__(256) k_ere_pairs(GT<GS> g, EreQueue Q, const u32* __restrict__ wl, const u32* __restrict__ wlCount) {
    const int clause_max = g.k.ere_clause_max;
    const u32 lane = LANE;
    const u32 groupsPerGrid = (gridDim.x * blockDim.x) / GS;
    const u32 count = *wlCount;
    __shared__ u32 sizeMask[8];   // bit s: some live clause has size s (255 = "255 or more"), k_ere_bloom
    if (threadIdx.x < 8) sizeMask[threadIdx.x] = g.bloom[(g.bloomMask >> 5) + 1 + threadIdx.x];
    __syncthreads();
    for (u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) / GS; wi < count; wi += groupsPerGrid) {
        const u32 v = g.elected[wl[wi]], p = V2L(v), n = p | 1u;
        const u32 ds = g.otSize[p], fs = g.otSize[n];
        if (!(ds && fs && ds <= g.k.ere_max_occurs && fs <= g.k.ere_max_occurs)) continue;
        const u32* P = g.occurs + g.otStart[p];
        const u32* N = g.occurs + g.otStart[n];
        // the reference tests the FIRST clause of the sorted lists, i.e. the smallest size of each list
        // (redundancy.cuh:151-154); the minimum does not need the lists sorted
        u32 minP = 0xFFFFFFFFu, minN = 0xFFFFFFFFu;
        for (u32 i = lane; i < ds; i += GS) minP = min(minP, g.hdr[P[i]].y);
        for (u32 j = lane; j < fs; j += GS) minN = min(minN, g.hdr[N[j]].y);
#pragma unroll
        for (int o = GS / 2; o; o >>= 1) { minP = min(minP, __shfl_xor_sync(FULL, minP, o, GS)); minN = min(minN, __shfl_xor_sync(FULL, minN, o, GS)); }
        if ((int)minP > clause_max || (int)minN > clause_max) continue;
        // pairs (i, j) in pos-major order, lane-strided; the indices advance without a division.  (A variant with the positive
        // clause uniform over the group and the negative headers cached in registers measured 15 % SLOWER on cfg2 -
        // profiles/r02_ab_c12.jsonl: lanes idle on the second pass of a 52-entry list, no parallelism over i.)
        u32 i = lane / fs, j = lane - i * fs;
        for (; i < ds; ereAdvance(i, j, (u32)GS, fs)) {
            const u32 ciP = P[i], ciN = N[j];
            const uint4 hp = g.hdr[ciP];
            if (C_DELETED(hp.w)) continue;
            const uint4 hn = g.hdr[ciN];
            if (C_DELETED(hn.w) || (int)(hp.y + hn.y - 2) > clause_max) continue;
            const u32* a = g.pool + hp.x; const int n1 = (int)hp.y;
            const u32* b = g.pool + hn.x; const int n2 = (int)hn.y;
            {   // cheapest filter first: bounds of the resolvent length from the two signatures alone.  A literal
                // of a can only be shared with b if its hash bit is set in b's signature, so at most U
                // literals merge and the length lies in [n1+n2-2-U, n1+n2-2]; no live clause of such a
                // size -> nothing can equal this resolvent (most pairs of a uniform k-SAT formula stop here)
                u32 U = 0;
                for (int q = 0; q < n1; q++) { const u32 l = a[q]; if (LABS(l) != v) U += (hn.z >> (l & 31u)) & 1u; }
                const u32 lenMax = (u32)(n1 + n2 - 2);
                if (U > (u32)(n2 - 1)) U = (u32)(n2 - 1);
                bool any = false;
                for (u32 sN = lenMax - U; sN <= lenMax; sN++) { const u32 sb = sN < 255u ? sN : 255u; any |= (sizeMask[sb >> 5] >> (sb & 31u)) & 1u; }
                if (!any) continue;
            }
            u32 len, first, last, sig;
            if (!ereKey(a, n1, b, n2, v, len, first, last, sig) || len <= 1) continue;
            // filters: is there a live clause of this size at all / with this key at all?
            { const u32 sb = len < 255u ? len : 255u; if (!((sizeMask[sb >> 5] >> (sb & 31u)) & 1u)) continue; }
            { const u32 hb = keyHash(len, first, last, sig) & g.bloomMask; if (!((g.bloom[hb >> 5] >> (hb & 31u)) & 1u)) continue; }
            u32 minsize;
            const u32 best = ereBest(g.otSize, a, n1, b, n2, v, minsize);
            if (!minsize) continue;
            if (Q.items) {
                const u32 slot = atomicAdd(Q.count, 1u);
                if (slot < Q.cap) { Q.items[3 * slot] = ciP; Q.items[3 * slot + 1] = ciN; Q.items[3 * slot + 2] = v; Q.need[best] = 1; }
                else *Q.overflow = 1u;
            } else {
                const u32 type = (C_LEARNT(hp.w) || C_LEARNT(hn.w)) ? CB_LEARNT : 0u;
                ereSearchDelete(g, a, n1, b, n2, v, len, first, last, sig, type, best, minsize);
            }
        }
    }
}
// Phase B: one thread per recorded resolvent, in queue order
__global__ void __launch_bounds__(256) k_ere_apply(GT<32> g, const u32* __restrict__ items, const u32* __restrict__ count) {
    const int clause_max = g.k.ere_clause_max;
    const u32 nq = *count;
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const u32 ciP = items[3 * q], ciN = items[3 * q + 1], v = items[3 * q + 2];
        const uint4 hp = g.hdr[ciP];
        if (C_DELETED(hp.w)) continue;
        const uint4 hn = g.hdr[ciN];
        if (C_DELETED(hn.w) || (int)(hp.y + hn.y - 2) > clause_max) continue;
        const u32* a = g.pool + hp.x; const int n1 = (int)hp.y;
        const u32* b = g.pool + hn.x; const int n2 = (int)hn.y;
        u32 len, first, last, sig, minsize;
        if (!ereKey(a, n1, b, n2, v, len, first, last, sig) || len <= 1) continue;
        const u32 best = ereBest(g.otSize, a, n1, b, n2, v, minsize);
        if (!minsize) continue;
        const u32 type = (C_LEARNT(hp.w) || C_LEARNT(hn.w)) ? CB_LEARNT : 0u;
        ereSearchDelete(g, a, n1, b, n2, v, len, first, last, sig, type, best, minsize);
    }
}

// ------------------------------------------------------------------ group-size classes
// elected[] indices split by occurrence count of the variable; order inside a class is arbitrary
// (nothing observable depends on it: scan offsets are indexed by the position in elected[])
__global__ void k_bin_reset(DevCounters* dc) { dc->bin[0] = dc->bin[1] = dc->bin[2] = dc->bin[3] = 0; }
__global__ void k_bin_elected(const u32* __restrict__ elected, const u32* nDev, u32 nHost, const u32* __restrict__ otSize, u32 t4, u32 t8,
                              u32* __restrict__ wl4, u32* __restrict__ wl8, u32* __restrict__ wl32, DevCounters* dc) {
    const u32 n = nDev ? *nDev : nHost;
    for (u32 i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const u32 i = i0 + threadIdx.x;
        int cls = -1;
        if (i < n) {
            const u32 x = elected[i];
            const u32 deg = otSize[V2L(x)] + otSize[V2L(x) | 1u];
            cls = deg <= t4 ? 0 : deg <= t8 ? 1 : 2;
        }
        for (int k = 0; k < 3; k++) {
            const u32 m = __ballot_sync(0xffffffffu, cls == k);
            if (!m) continue;
            u32 base = 0;
            const u32 leader = __ffs(m) - 1;
            if (laneId() == leader) base = atomicAdd(&dc->bin[k], __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (cls == k) (k == 0 ? wl4 : k == 1 ? wl8 : wl32)[base + __popc(m & lanemaskLt())] = i;
        }
    }
}

// ------------------------------------------------------------------ kernel profile mode: algorithmic bytes of the per-variable kernels
// SUB / BVE / BCE / ERE must read, for every variable they are given, every clause of its two occurrence
// lists: list entry (4) + header (16) + literals (4|c|).  Summed on the device over a worklist; `filter` (BVE phase 3)
// keeps the variables with a recorded elimination type.
__global__ void __launch_bounds__(256) k_gather_bytes(const u32* __restrict__ wl, const u32* __restrict__ count, u32 upper, const u32* __restrict__ elected,
                                                      const u32* __restrict__ otStart, const u32* __restrict__ otSize, const u32* __restrict__ occurs,
                                                      const uint4* __restrict__ hdr, const u32* __restrict__ filter, unsigned long long* out) {
    const u32 n = count ? *count : upper;
    unsigned long long b = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 item = wl ? wl[i] : i;
        if (filter && !RECOVERTYPE(filter[item])) continue;
        const u32 x = elected ? elected[item] : (item & 0x7FFFFFFFu);
        for (u32 side = 0; side < 2; side++) {
            const u32 lit = V2L(x) | side;
            const u32 m = otSize[lit];
            const u32* list = occurs + otStart[lit];
            for (u32 j = 0; j < m; j++) b += 20u + 4u * hdr[list[j]].y;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if ((threadIdx.x & 31u) == 0 && b) atomicAdd(out, b);
}
static double gatherBytes(Ctx* c, const u32* wl, const u32* countDev, u32 upper, bool direct, const u32* filter) {
    if (!c->ktOn || !upper) return 0.0;
    unsigned long long* out = (unsigned long long*)&c->dc->profBytes;
    unsigned long long b = 0;
    cudaMemsetAsync(out, 0, 8, c->stream);
    k_gather_bytes<<<gridFor(upper, 256), 256, 0, c->stream>>>(wl, countDev, upper, direct ? nullptr : c->elected, c->otStart, c->otSize, c->occurs,
                                                             c->hdr[c->cur], filter, out);
    cudaMemcpyAsync(&b, out, 8, cudaMemcpyDeviceToHost, c->stream);
    cudaMemsetAsync(out, 0, 8, c->stream);
    cudaStreamSynchronize(c->stream);
    return (double)b;
}
double profGather(Ctx* c, const u32* wl, const u32* countDev, u32 upper, bool direct) { return gatherBytes(c, wl, countDev, upper, direct, nullptr); }
// bytes of the three group-size classes (worklists of k_bin_elected), taken BEFORE the stage runs (it shrinks the lists)
struct ClassBytes { double b[3]; };
static ClassBytes classBytes(Ctx* c, const u32* filter = nullptr) {
    ClassBytes cb = {{0, 0, 0}};
    if (!c->ktOn) return cb;
    cb.b[0] = gatherBytes(c, c->wlA, &c->dc->bin[0], c->numElected, false, filter);
    cb.b[1] = gatherBytes(c, c->wlB, &c->dc->bin[1], c->numElected, false, filter);
    cb.b[2] = gatherBytes(c, c->sortK, &c->dc->bin[2], c->numElected, false, filter);
    return cb;
}

// ------------------------------------------------------------------ host launchers
static G makeG(Ctx* c, const KOpts& k) {
    G g;
    g.hdr = c->hdr[c->cur]; g.pool = c->pool[c->cur];
    g.otStart = c->otStart; g.otSize = c->otSize; g.occurs = c->occurs; g.key = c->key;
    g.elected = c->elected; g.eliminated = c->eliminated; g.vorg = c->vorg; g.varcore = c->varcore;
    g.units = c->units; g.unitsCap = 2 * (c->V + 1); g.resolved = c->resolved; g.resolvedCap = c->resolvedCap;
    g.veType = c->veType; g.veUcnt = c->veUcnt; g.veRpos = c->veRpos; g.veRref = c->veRref;
    g.dc = c->dc; g.k = k; g.numElected = c->numElected;
    g.bloom = nullptr; g.bloomMask = 0;
    return g;
}
template <int GS> static GT<GS> asGroup(const G& g) { GT<GS> t; static_cast<G&>(t) = g; return t; }
// enough CTAs for `n` groups of GS lanes, capped: the kernels stride over their worklist
static u32 groupGrid(u32 nGroups, u32 GS, u32 block) {
    const u64 b = ((u64)nGroups * GS + block - 1) / block;
    const u64 cap = 148ull * 16;
    return (u32)(b > cap ? cap : (b ? b : 1));
}
// ------------------------------------------------------------------ shape order inside a class
// A 4- or 8-lane group shares its warp with 7 or 3 other variables; what a variable does in SUB/BVE
// (which gate pattern matches, how many pairs are merged) follows from the shape of its two lists,
// so the worklist of a class is grouped by (|pos|, |neg|): warps run one code path instead of
// eight (the BVE kernel is ~270 KB of SASS - divergent groups also thrash the instruction cache).
// Counting sort with 256 keys, tile-local ranks + one global atomic per (tile, key); the order inside
// a key is arbitrary, which nothing observes (results are indexed by the position in elected[]).
#define SHAPE_TILE 2048
__device__ __forceinline__ u32 shapeKey(const u32* __restrict__ elected, const u32* __restrict__ otSize, u32 item) {
    const u32 x = elected[item];
    const u32 np = otSize[V2L(x)], nn = otSize[V2L(x) | 1u];
    return (min(np, 15u) << 4) | min(nn, 15u);
}
__global__ void __launch_bounds__(256) k_shape_hist(const u32* __restrict__ wl, const u32* __restrict__ count, const u32* __restrict__ elected,
                                                    const u32* __restrict__ otSize, u32* __restrict__ hist) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u32 n = *count;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&h[shapeKey(elected, otSize, wl[i])], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_shape_scan(u32* __restrict__ hist) {   // hist[0..255] -> exclusive starts in hist[256..511]
    __shared__ u32 s[256];
    s[threadIdx.x] = hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) { u32 run = 0; for (int k = 0; k < 256; k++) { const u32 t = s[k]; s[k] = run; run += t; } }
    __syncthreads();
    hist[256 + threadIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(256) k_shape_scatter(const u32* __restrict__ wl, const u32* __restrict__ count, const u32* __restrict__ elected,
                                                       const u32* __restrict__ otSize, u32* __restrict__ cursor, u32* __restrict__ out) {
    __shared__ u32 cnt[256], base[256];
    const u32 n = *count;
    for (u32 t0 = blockIdx.x * SHAPE_TILE; t0 < n; t0 += gridDim.x * SHAPE_TILE) {
        __syncthreads();
        cnt[threadIdx.x] = 0;
        __syncthreads();
        u32 item[SHAPE_TILE / 256], key[SHAPE_TILE / 256], rank[SHAPE_TILE / 256];
#pragma unroll
        for (int k = 0; k < SHAPE_TILE / 256; k++) {
            const u32 i = t0 + k * 256 + threadIdx.x;
            key[k] = 0xFFFFFFFFu;
            if (i < n) { item[k] = wl[i]; key[k] = shapeKey(elected, otSize, item[k]); rank[k] = atomicAdd(&cnt[key[k]], 1u); }
        }
        __syncthreads();
        base[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], cnt[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SHAPE_TILE / 256; k++) if (key[k] != 0xFFFFFFFFu) out[base[key[k]] + rank[k]] = item[k];
    }
}
static void shapeSort(Ctx* c, u32* wl, const u32* count, u32 upperBound) {
    if (upperBound < (1u << 16)) return;   // small worklists: four more launches per class cost more than the divergence they remove (cfg1: -0.3 ms)
    u32* hist = c->radixHist;      // 512 words, free after the election's radix sort
    u32* tmp = c->flagA;           // free until BVE phase 3 writes the survivor flags
    cudaMemsetAsync(hist, 0, 256 * 4, c->stream);
    LAUNCH(c, k_shape_hist, gridFor(upperBound, 256, 4), 256, 0, wl, count, c->elected, c->otSize, hist);
    LAUNCH(c, k_shape_scan, 1, 256, 0, hist);
    LAUNCH(c, k_shape_scatter, gridFor(upperBound, 256, SHAPE_TILE / 256), 256, 0, wl, count, c->elected, c->otSize, hist + 256, tmp);
    LAUNCH(c, k_copy_u32, gridFor(upperBound, 256), 256, 0, tmp, wl, count);
}

// worklists of the three classes live in the MIS scratch (free between election and the next round)
static void binElected(Ctx* c, const KOpts& k, bool countOnDevice, bool byShape = false) {
    // long XOR arities need the big shared-memory slice: everything runs on full warps then
    const u32 t4 = k.xor_max_arity + 2 > VE_SLICE_SMALL ? 0 : BIN_T4, t8 = k.xor_max_arity + 2 > VE_SLICE_SMALL ? 0 : BIN_T8;
    LAUNCH(c, k_bin_reset, 1, 1, 0, c->dc);
    LAUNCH(c, k_bin_elected, gridFor(c->numElected, 256), 256, 0, c->elected, countOnDevice ? &c->dc->numElected : nullptr, c->numElected,
           c->otSize, t4, t8, c->wlA, c->wlB, c->sortK, c->dc);
    if (byShape) {
        shapeSort(c, c->wlA, &c->dc->bin[0], c->numElected);
        shapeSort(c, c->wlB, &c->dc->bin[1], c->numElected);
    }
}

// ------------------------------------------------------------------ device DRAT stream (proof.cu, proofutils.cuh)
// The reference threads proof code through every elimination kernel (count the bytes, reserve with
// cuVecB::jump, write).  Here the hot kernels stay as they are and the lines are produced by small passes
// that run only with opts.proof_en, right after the stage whose effect they record:
//   SUB   k_proof_stream(MOLTEN) : a strengthened clause still carries its molten mark (k_sub keeps it in proof
//                                  mode) -> 'a' line with the new literals, mark cleared (saveProof, proofutils.cuh:182-200)
//         k_proof_stream(NEWDEL) : clauses deleted since the k_proof_snap bitmap -> 'd' lines
//   BVE   k_proof_equ            : clauses rewritten by an equivalence substitution -> 'a' (addProof, equivalence.cuh:100-109)
//         k_proof_units, k_proof_stream(RANGE) : the resolvents of phase 3, units and clauses -> 'a' (bounded.cuh:76-120)
//   BCE / ERE  k_proof_snap + k_proof_stream(NEWDEL) -> 'd' (blocked.cuh:67-72, redundancy.cuh:122-129)
// All of them stream headers (16 B per clause) and touch literals only of the clauses they print.  A warp sizes its
// 32 lines, reserves their bytes with ONE atomic on dc->proofSize and writes them in order.
struct ProofOut { unsigned char* buf; };
#define PROOF_MOLTEN 1u
#define PROOF_NEWDEL 2u
#define PROOF_RANGE 3u
// bytes of the 7-bit varint of the ORIGINAL literal (BLUT / COUNTBYTES, proof.cu:31-41)
__device__ __forceinline__ u32 proofOrgLit(const u32* __restrict__ vorg, u32 lit) { return (vorg[LABS(lit)] << 1) | LSIGN(lit); }
__device__ __forceinline__ u32 proofLitBytes(const u32* __restrict__ vorg, u32 lit) { return (38u - (u32)__clz((int)proofOrgLit(vorg, lit))) / 7u; }
__device__ __forceinline__ u32 proofClauseBytes(const u32* __restrict__ vorg, const u32* lits, u32 n) {
    u32 b = 2;   // prefix + terminating 0 (countProofBytes, proofutils.cuh:34-60)
    for (u32 k = 0; k < n; k++) b += proofLitBytes(vorg, lits[k]);
    return b;
}
// saveProofClause (proofutils.cuh:123-170)
__device__ __forceinline__ void proofWriteClause(unsigned char* out, const u32* __restrict__ vorg, const u32* lits, u32 n, unsigned char state) {
    *out++ = state;
    for (u32 k = 0; k < n; k++) {
        u32 org = proofOrgLit(vorg, lits[k]);
        while (org & 0xFFFFFF80u) { *out++ = (unsigned char)((org & 0x7Fu) | 0x80u); org >>= 7; }
        *out++ = (unsigned char)org;
    }
    *out = 0;
}
// the 32 lanes of a warp append their lines (bytes == 0: none) in lane order; returns this lane's write pointer or null
__device__ __forceinline__ unsigned char* proofReserve(const G& g, const ProofOut& po, u32 bytes) {
    const u32 incl = warpIncl(bytes);
    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
    if (!total) return nullptr;
    u32 base = 0;
    if (laneId() == 0) {
        base = atomicAdd(&g.dc->proofSize, total);
        if ((u64)base + total > g.dc->proofCap) atomicOr(&g.dc->flags, 32u);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!bytes || (u64)base + total > g.dc->proofCap) return nullptr;
    return po.buf + base + (incl - bytes);
}
// snap bit i = clause i is deleted now
__global__ void __launch_bounds__(256) k_proof_snap(const uint4* __restrict__ hdr, u32 n, u32* __restrict__ snap) {
    const u32 nWarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5; base < n; base += nWarps << 5) {
        const u32 i = base + laneId();
        const u32 m = __ballot_sync(0xffffffffu, i < n && C_DELETED(hdr[i].w));
        if (laneId() == 0) snap[base >> 5] = m;
    }
}
__global__ void __launch_bounds__(256) k_proof_stream(G g, ProofOut po, const u32* __restrict__ snap, u32 lo, u32 hiHost, const u32* hiDev, u32 mode) {
    const u32 hi = hiDev ? *hiDev : hiHost;
    const u32 nWarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 base = lo + (((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5); base < hi; base += nWarps << 5) {
        const u32 i = base + laneId();
        uint4 h = make_uint4(0, 0, 0, 0);
        bool emit = false;
        if (i < hi) {
            h = g.hdr[i];
            // a clause deleted before the snapshot may still carry the molten mark of the gate search that preceded its
            // elimination (deleteAll keeps it); only clauses alive when SUB started can have been strengthened by it
            if (mode == PROOF_MOLTEN) emit = C_MOLTEN(h.w) && !((snap[i >> 5] >> (i & 31u)) & 1u);
            else if (mode == PROOF_NEWDEL) emit = C_DELETED(h.w) && !((snap[i >> 5] >> (i & 31u)) & 1u);
            else emit = !C_DELETED(h.w) && h.y > 0;
        }
        const u32 bytes = emit ? proofClauseBytes(g.vorg, g.pool + h.x, h.y) : 0u;
        unsigned char* out = proofReserve(g, po, bytes);
        if (out) proofWriteClause(out, g.vorg, g.pool + h.x, h.y, mode == PROOF_NEWDEL ? (unsigned char)'d' : (unsigned char)'a');
        if (emit && mode == PROOF_MOLTEN) g.hdr[i].w = h.w & ~CB_MOLTEN;   // c.freeze() (proofutils.cuh:191)
    }
}
// resolvent units of phase 3: units[dc->proofUnits0 .. dc->numUnits) (saveProofUnit, proofutils.cuh:123-128)
__global__ void k_proof_mark_units(DevCounters* dc) { dc->proofUnits0 = dc->numUnits; }
__global__ void __launch_bounds__(256) k_proof_units(G g, ProofOut po) {
    const u32 lo = g.dc->proofUnits0, hi = min(g.dc->numUnits, g.unitsCap);
    const u32 nWarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 base = lo + (((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5); base < hi; base += nWarps << 5) {
        const u32 i = base + laneId();
        u32 lit = 0;
        if (i < hi) lit = g.units[i];
        const u32 bytes = lit ? 2u + proofLitBytes(g.vorg, lit) : 0u;
        unsigned char* out = proofReserve(g, po, bytes);
        if (out) proofWriteClause(out, g.vorg, &lit, 1, (unsigned char)'a');
    }
}
// Equivalence substitution: a variable eliminated in phase 1 without resolvents keeps its lists only if it was
// substituted (toblivion clears them), and what is still ORIGINAL in them are the rewritten clauses - negative list
// first (equivalence.cuh:107-108).  One warp per elected variable, before the survivors are compacted.
__global__ void __launch_bounds__(256) k_proof_equ(G g, ProofOut po, u32 E) {
    const u32 nWarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 tid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tid < E; tid += nWarps) {
        const u32 x = g.elected[tid];
        const u32 e = g.eliminated[x];
        if (!(e & MELTING_MASK) || (e & ADDING_MASK)) continue;
        for (int side = 1; side >= 0; side--) {
            const u32 lit = V2L(x) | (u32)side;
            const u32 cnt = g.otSize[lit];
            const u32* list = g.occurs + g.otStart[lit];
            for (u32 base = 0; base < cnt; base += 32) {
                const u32 j = base + laneId();
                uint4 h = make_uint4(0, 0, 0, 0);
                bool emit = false;
                if (j < cnt) { h = g.hdr[list[j]]; emit = C_ORIGINAL(h.w) && h.y > 0; }
                const u32 bytes = emit ? proofClauseBytes(g.vorg, g.pool + h.x, h.y) : 0u;
                unsigned char* out = proofReserve(g, po, bytes);
                if (out) proofWriteClause(out, g.vorg, g.pool + h.x, h.y, (unsigned char)'a');
            }
        }
    }
}
// The counting functions of the reference refuse a candidate whose resolvents need more than ADDEDPROOF_MAX proof bytes
// or produce more than ADDEDCLS_MAX units (resolve.cuh:66-70, :152-154, elimination.cuh:445-447, function.cuh:232-235).
// Phase 1 does not count proof bytes; a candidate that could reach the limit (bound from its literal count and the
// widest literal) is reported instead of being decided differently - it takes tens of thousands of added literals.
__global__ void __launch_bounds__(256) k_proof_guard(G g, u32 E, u32 bmax) {
    for (u32 tid = blockIdx.x * blockDim.x + threadIdx.x; tid < E; tid += gridDim.x * blockDim.x) {
        const u32 info = g.veType[tid];
        if (!RECOVERTYPE(info)) continue;
        const u64 cls = RECOVERADDEDCLS(info), lits = RECOVERADDEDLITS(info), units = g.veUcnt[tid];
        if (units > ADDEDCLS_MAX || (u64)bmax * lits + 2 * cls + (u64)(bmax + 2) * units > 0x3FFFFull) atomicOr(&g.dc->flags, 16u);
    }
}
// cuPROOF::count (proof.cu:65-121): proof bytes of every literal of the loaded formula
__global__ void __launch_bounds__(256) k_proof_count(const u32* __restrict__ lits, u64 n, const u32* __restrict__ vorg, DevCounters* dc) {
    u32 local = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) local += proofLitBytes(vorg, lits[i]);
    local = warpSum(local);
    if (laneId() == 0 && local) atomicAdd((unsigned long long*)&dc->proofLitBytes, (unsigned long long)local);
}
__global__ void k_proof_cap(DevCounters* dc) { dc->proofCap = (u32)((double)(u32)dc->proofLitBytes * 1.5); dc->proofSize = 0; }   // simplify.cu:130-131
void launchProofCount(Ctx* c) {
    LAUNCH(c, k_proof_count, gridFor(c->L0, 256, 8), 256, 0, c->inLits, c->L0, c->vorg, c->dc);
    LAUNCH(c, k_proof_cap, 1, 1, 0, c->dc);
}
static void proofSnap(Ctx* c) {
    const u32 n = c->hdc->numCls;
    LAUNCH(c, k_proof_snap, gridFor(n, 256), 256, 0, c->hdr[c->cur], n, c->proofSnap);
}
static void proofStream(Ctx* c, const G& g, u32 lo, u32 hiHost, const u32* hiDev, u32 mode, u64 sizeHint) {
    ProofOut po{c->proofBuf};
    LAUNCH(c, k_proof_stream, gridFor(sizeHint, 256), 256, 0, g, po, c->proofSnap, lo, hiHost, hiDev, mode);
}
#define LAUNCH_CLASSES(c, kern, block, E, g, cb, ...)                                                        \
    do {                                                                                                     \
        LAUNCH(c, kern<4>, groupGrid(E, 4, block), block, 0, asGroup<4>(g), ##__VA_ARGS__, c->wlA, &c->dc->bin[0]);   \
        KB(c, (cb).b[0]);                                                                                    \
        LAUNCH(c, kern<8>, groupGrid(E, 8, block), block, 0, asGroup<8>(g), ##__VA_ARGS__, c->wlB, &c->dc->bin[1]);   \
        KB(c, (cb).b[1]);                                                                                    \
        LAUNCH(c, kern<32>, groupGrid(E, 32, block), block, 0, asGroup<32>(g), ##__VA_ARGS__, c->sortK, &c->dc->bin[2]); \
        KB(c, (cb).b[2]);                                                                                    \
    } while (0)

void launchSUB(Ctx* c, const KOpts& k) {
    if (!c->numElected) return;
    G g = makeG(c, k);
    binElected(c, k, false, true);
    if (k.proof_en) proofSnap(c);
    const ClassBytes cb = classBytes(c);
    // the local-copy kernels win on short clauses (3-SAT, Tseitin gates: cfg1 -15 %, cfg3 -13 %) and lose on the 4-literal
    // adder clauses of cfg4 (+50 % on the 8-lane class: fixed 8-word slots, fewer resident CTAs) - profiles/r02_ab_local_c3.jsonl
    static const int subEnv = getenv("SIGMA_SUB_LOCAL") ? atoi(getenv("SIGMA_SUB_LOCAL")) : -1;   // 0 / 1: force (A/B measurements)
    const bool subLocal = subEnv >= 0 ? subEnv != 0 : c->L0 * 10 <= c->C0 * 31;   // decided once per call, from the loaded formula
    if (subLocal) {
        const u32 E = c->numElected;
        u32* redo = c->rank;       // rank[] is dead after the election
        u32* redoCount = &c->dc->bin[3];
        LAUNCH(c, k_sub_local<4>, groupGrid(E, 4, 128), 128, 0, asGroup<4>(g), c->wlA, &c->dc->bin[0], redo, redoCount);
        KB(c, cb.b[0]);
        LAUNCH(c, k_sub_local<8>, groupGrid(E, 8, 128), 128, 0, asGroup<8>(g), c->wlB, &c->dc->bin[1], redo, redoCount);
        KB(c, cb.b[1]);
        LAUNCH(c, k_sub<32>, groupGrid(E, 32, 128), 128, 0, asGroup<32>(g), c->sortK, &c->dc->bin[2]);
        KB(c, cb.b[2]);
        LAUNCH(c, k_sub<32>, 148, 128, 0, asGroup<32>(g), redo, redoCount);   // variables with a long clause, handed over by the small groups
    } else
        LAUNCH_CLASSES(c, k_sub, 128, c->numElected, g, cb);
    if (k.proof_en) {   // subsume.cuh:465-475: strengthened clauses added, then subsumed ones deleted
        const u32 n = c->hdc->numCls;
        proofStream(c, g, 0, n, nullptr, PROOF_MOLTEN, n);
        proofStream(c, g, 0, n, nullptr, PROOF_NEWDEL, n);
    }
}

// veAsync + postVE (elimination.cu:131-154, 235-266)
void launchVE(Ctx* c, const KOpts& k) {
    if (!c->numElected) return;
    G g = makeG(c, k);
    const u32 E = c->numElected;
    LAUNCH(c, k_ve_reset, 1, 1, 0, c->dc);
    binElected(c, k, false, true);   // SUB shrank the lists: classes by the current sizes
    u32* redo = c->rank;       // rank[] is dead after the election
    u32* redoCount = &c->dc->bin[3];
    const ClassBytes cb1 = classBytes(c);   // + 20 bytes per variable: type, ucnt, rpos, rref (SURVEY 8d "BVE count")
    static const int veEnv = getenv("SIGMA_VE_LOCAL") ? atoi(getenv("SIGMA_VE_LOCAL")) : -1;   // 0 / 1: force (A/B measurements)
    const bool veLocal = veEnv >= 0 ? veEnv != 0 : c->L0 * 10 <= c->C0 * 31;   // see launchSUB
    if (veLocal) {
        LAUNCH(c, k_ve_phase1_local<4>, groupGrid(E, 4, 128), 128, 0, asGroup<4>(g), c->wlA, &c->dc->bin[0], redo, redoCount);
        KB(c, cb1.b[0]);
        LAUNCH(c, k_ve_phase1_local<8>, groupGrid(E, 8, 128), 128, 0, asGroup<8>(g), c->wlB, &c->dc->bin[1], redo, redoCount);
        KB(c, cb1.b[1]);
    } else {
        LAUNCH(c, k_ve_phase1<4>, groupGrid(E, 4, 128), 128, 0, asGroup<4>(g), c->wlA, &c->dc->bin[0], redo, redoCount);
        KB(c, cb1.b[0]);
        LAUNCH(c, k_ve_phase1<8>, groupGrid(E, 8, 128), 128, 0, asGroup<8>(g), c->wlB, &c->dc->bin[1], redo, redoCount);
        KB(c, cb1.b[1]);
    }
    LAUNCH(c, k_ve_phase1<32>, groupGrid(E, 32, 128), 128, 0, asGroup<32>(g), c->sortK, &c->dc->bin[2], redo, redoCount);
    KB(c, cb1.b[2] + 20.0 * E);
    LAUNCH(c, k_ve_phase1<32>, 148, 128, 0, asGroup<32>(g), redo, redoCount, redo, redoCount);   // variables handed over by the small groups
    if (k.proof_en) {
        LAUNCH(c, k_proof_guard, gridFor(E, 256), 256, 0, g, E, c->proofBMax);
        LAUNCH(c, k_proof_mark_units, 1, 1, 0, c->dc);   // units before this mark come from substitutions (no unit lines)
    }
    // phase 2: exclusive scans seeded with the current CNF sizes
    VEBase vb;
    vb.numCls0 = c->hdc->numCls; vb.poolUsed0 = c->hdc->poolUsed; vb.dataSize0 = c->hdc->dataSize;
    scanExclusiveU32(c, c->veRpos, c->veRpos, E, vb.numCls0, nullptr);
    scanExclusiveU64(c, c->veRref, c->veRref, E, vb.dataSize0);
    // BVE emit: the lists of the variables that resolve + (kernel profile mode, after the fact) the resolvents written
    const ClassBytes cb3 = classBytes(c, c->veType);
    int kid3[3] = {-1, -1, -1};
    LAUNCH(c, k_ve_phase3<4>, groupGrid(E, 4, 128), 128, 0, asGroup<4>(g), vb, c->flagA, c->wlA, &c->dc->bin[0]);
    KB(c, cb3.b[0]); kid3[0] = c->ktLastId;
    LAUNCH(c, k_ve_phase3<8>, groupGrid(E, 8, 128), 128, 0, asGroup<8>(g), vb, c->flagA, c->wlB, &c->dc->bin[1]);
    KB(c, cb3.b[1]); kid3[1] = c->ktLastId;
    LAUNCH(c, k_ve_phase3<32>, groupGrid(E, 32, 128), 128, 0, asGroup<32>(g), vb, c->flagA, c->sortK, &c->dc->bin[2]);
    KB(c, cb3.b[2]); kid3[2] = c->ktLastId;
    LAUNCH(c, k_ve_resize, 1, 1, 0, g, vb);
    if (c->ktOn && kid3[0] >= 0) {   // resolvents written: 16 bytes of header + 4 per literal; shared out by the classes' read bytes
        DevCounters t;
        cudaMemcpyAsync(&t, c->dc, sizeof t, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        const double wr = 16.0 * (t.numCls - vb.numCls0) + 4.0 * (t.poolUsed - vb.poolUsed0);
        const double tot = cb3.b[0] + cb3.b[1] + cb3.b[2];
        for (int q = 0; q < 3; q++) if (tot > 0) c->ktBytes[kid3[q]] += wr * cb3.b[q] / tot;
    }
    if (k.proof_en) {   // before elected[] is compacted: k_proof_equ walks the lists of the variables eliminated in phase 1
        ProofOut po{c->proofBuf};
        LAUNCH(c, k_proof_equ, gridFor((u64)E * 32, 256), 256, 0, g, po, E);
        LAUNCH(c, k_proof_units, 148, 256, 0, g, po);
        proofStream(c, g, vb.numCls0, 0, &c->dc->numCls, PROOF_RANGE, c->capC - vb.numCls0);
    }
    // elected := survivors, order kept (cub::DeviceSelect::If in postVE)
    scanExclusiveU32(c, c->flagA, c->flagB, E, 0, &c->dc->numElected);
    {
        // flagA still holds the flags, flagB the positions; compact through sortV then copy back
        LAUNCH(c, k_select_scatter, gridFor(E, 256), 256, 0, c->elected, c->flagA, c->flagB, E, c->sortV);
        LAUNCH(c, k_copy_u32, gridFor(E, 256), 256, 0, c->sortV, c->elected, &c->dc->numElected);
    }
}

void launchBCE(Ctx* c, const KOpts& k) {
    if (!c->numElected) return;
    G g = makeG(c, k);
    binElected(c, k, false);
    if (k.proof_en) proofSnap(c);
    const ClassBytes cb = classBytes(c);
    LAUNCH_CLASSES(c, k_bce, 128, c->numElected, g, cb);
    if (k.proof_en) proofStream(c, g, 0, c->hdc->numCls, nullptr, PROOF_NEWDEL, c->hdc->numCls);   // blocked.cuh:67-72
}

void launchERE(Ctx* c, const KOpts& k) {
    if (!c->numElected) return;
    G g = makeG(c, k);
    // Bloom filter over the keys of the live clauses, in the partition buffer of the OT build (free now)
    const u32 n = c->hdc->numCls;
    const u64 bufBytes = ((u64)c->capW + 4) * sizeof(uint2);
    u64 bits = 256;
    while (bits < 8ull * n && bits < (1ull << 30)) bits <<= 1;   // >= 8 bits per clause: the filter stays L2 resident
    while (bits > 32 && bits / 8 + 32 > bufBytes / 2) bits >>= 1;
    u32* bloom = (u32*)c->otPairs;
    cudaMemsetAsync(bloom, 0, bits / 8 + 32, c->stream);   // + the 256-bit clause-size mask
    LAUNCH(c, k_ere_bloom, gridFor(n, 256), 256, 0, c->key, n, bloom, (u32)(bits - 1));
    KB(c, 16.0 * n);   // one key per clause slot (the filter bits stay in L2)
    g.bloom = bloom; g.bloomMask = (u32)(bits - 1);
    binElected(c, k, false);
    // Phase A records the resolvents that pass the filters; only the lists they will be searched in
    // get sorted (sortOT, segsort.cu:37-48, restricted), then phase B searches and deletes.
    EreQueue Q;
    const u64 qOff = (bits / 8 + 32 + 255) & ~255ull;
    Q.items = (u32*)((char*)c->otPairs + qOff);
    const u64 qCap = (bufBytes - qOff) / 12;
    Q.cap = (u32)(qCap > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : qCap);
    if (const char* e = getenv("SIGMA_ERE_QUEUE_CAP")) { const u32 v = (u32)atoi(e); if (v < Q.cap) Q.cap = v; }   // tests: force the overflow fallback
    Q.count = &c->dc->scratch[3]; Q.overflow = &c->dc->scratch[4]; Q.need = c->needSort;
    cudaMemsetAsync(Q.count, 0, 8, c->stream);
    cudaMemsetAsync(c->needSort, 0, c->ND, c->stream);
    if (k.proof_en) proofSnap(c);   // the first pass only records: nothing is deleted before the snapshot is taken
    const ClassBytes cb = classBytes(c);
    LAUNCH_CLASSES(c, k_ere_pairs, 256, c->numElected, g, cb, Q);
    if (syncCounters(c)) return;
    const u32 nq = c->hdc->scratch[3];
    if (c->hdc->scratch[4]) {   // more survivors than the queue holds: sort everything, search in place (nothing was deleted yet)
        launchSortOT(c, 0);
        Q.items = nullptr;
        LAUNCH_CLASSES(c, k_ere_pairs, 256, c->numElected, g, cb, Q);
    } else if (nq) {
        launchSortOT(c, 2);
        LAUNCH(c, k_ere_apply, gridFor(nq, 256), 256, 0, asGroup<32>(g), Q.items, Q.count);
        KB(c, 12.0 * nq + 2.0 * 36.0 * nq + 10.0 * 20.0 * nq);   // queue entry, the two parents, ~10 key probes of the binary search
    }
    if (k.proof_en) proofStream(c, g, 0, n, nullptr, PROOF_NEWDEL, n);   // redundancy.cuh:122-129
}
