// parafrost_b200/csrc/prop.cu -- unit propagation of the units produced by SUB/BVE, then clause
// clean-up.  Replaces Solver::prop (src/gpu/elimbcp.cu:144-215) and its kernels
// bcp_seed_k / bcp_propagate_k / bcp_advance_k / bcp_apply_k (elimbcp.cu:43-142).
//
// Level-synchronous BFS over falsified literals; one warp per frontier literal, lanes stride
// over its occurrence list.  The closure is confluent, so the resulting assignment, the
// cleaned clauses and the set of forced variables equal the reference's; only the order of
// the derived units on the trail is scheduling dependent (there as here).
#include "common.cuh"

#define BCP_UNSET 0u
#define BCP_TRUE 1u
#define BCP_FALSE 2u

__device__ __forceinline__ u32 bcpLitVal(const u32* state, u32 lit) {
    const u32 s = state[LABS(lit)];
    if (s == BCP_UNSET) return BCP_UNSET;
    const bool sat = LSIGN(lit) ? (s == BCP_FALSE) : (s == BCP_TRUE);
    return sat ? BCP_TRUE : BCP_FALSE;
}

__global__ void k_bcp_seed(u32* state, u32* __restrict__ front, DevCounters* dc, unsigned char* __restrict__ eliminated,
                           const u32* __restrict__ units, u32 nunits) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nunits; i += gridDim.x * blockDim.x) {
        const u32 u = units[i], v = LABS(u);
        const u32 desired = LSIGN(u) ? BCP_FALSE : BCP_TRUE;
        const u32 old = atomicCAS(&state[v], BCP_UNSET, desired);
        if (old == BCP_UNSET) {
            eliminated[v] |= FORCED_MASK;
            atomicAdd(&dc->unassignedDec, 1u);   // variables newly assigned: a unit emitted twice counts once
            front[atomicAdd(&dc->bcpCurr, 1u)] = LFLIP(u);
        } else if (old != desired) dc->bcpConfl = 1;
    }
}

__global__ void __launch_bounds__(256) k_bcp_level(const uint4* __restrict__ hdr, const u32* __restrict__ pool,
                                                   const u32* __restrict__ otStart, const u32* __restrict__ otSize,
                                                   const u32* __restrict__ occurs, u32* state,
                                                   unsigned char* __restrict__ eliminated, u32* frontA, u32* frontB,
                                                   DevCounters* dc, u32* __restrict__ units, u32 unitsCap) {
    if (dc->bcpConfl) return;
    const u32 level = dc->bcpLevel;
    const u32* frontCur = (level & 1u) ? frontB : frontA;
    u32* frontNext = (level & 1u) ? frontA : frontB;
    const u32 curSize = dc->bcpCurr;
    const u32 lane = threadIdx.x & 31u;
    const u32 warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    for (u32 fi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; fi < curSize; fi += warpsPerGrid) {
        const u32 ulit = frontCur[fi];
        const u32 n = otSize[ulit];
        const u32* list = occurs + otStart[ulit];
        for (u32 j = lane; j < n; j += 32) {
            const uint4 h = hdr[list[j]];
            if (C_DELETED(h.w)) continue;
            const u32* l = pool + h.x;
            u32 unit = 0; int nunset = 0; bool sat = false;
            for (u32 k = 0; k < h.y; k++) {
                const u32 ve = bcpLitVal(state, l[k]);
                if (ve == BCP_TRUE) { sat = true; break; }
                if (ve == BCP_UNSET) { unit = l[k]; if (++nunset > 1) break; }
            }
            if (sat) continue;
            if (!nunset) dc->bcpConfl = 1;
            else if (nunset == 1) {
                const u32 v = LABS(unit);
                const u32 desired = LSIGN(unit) ? BCP_FALSE : BCP_TRUE;
                const u32 old = atomicCAS(&state[v], BCP_UNSET, desired);
                if (old == BCP_UNSET) {
                    eliminated[v] |= FORCED_MASK;
                    atomicAdd(&dc->unassignedDec, 1u);
                    const u32 slot = atomicAdd(&dc->numUnits, 1u);
                    if (slot < unitsCap) units[slot] = unit; else dc->flags |= 2u;
                    frontNext[atomicAdd(&dc->bcpNext, 1u)] = LFLIP(unit);
                } else if (old != desired) dc->bcpConfl = 1;
            }
        }
    }
}

__global__ void k_bcp_advance(DevCounters* dc) {
    dc->bcpCurr = dc->bcpNext;
    dc->bcpNext = 0;
    dc->bcpLevel++;
}

// bcp_apply_k (elimbcp.cu:121-142): delete satisfied clauses, strip falsified literals, new signature
__global__ void k_bcp_apply(uint4* __restrict__ hdr, u32* __restrict__ pool, u32 n, const u32* __restrict__ state) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        u32* l = pool + h.x;
        u32 sig = 0, newsz = 0; bool sat = false;
        for (u32 k = 0; k < h.y; k++) {
            const u32 lit = l[k];
            const u32 ve = bcpLitVal(state, lit);
            if (ve == BCP_TRUE) { sat = true; break; }
            if (ve == BCP_FALSE) continue;
            l[newsz++] = lit; sig |= MAPHASH(lit);
        }
        if (sat) h.w = (h.w & ~CB_ST_MASK) | CB_DELETED;
        else { h.z = sig; h.y = newsz; }
        hdr[i] = h;
    }
}

// host enqueue loop (elimbcp.cu:186-201): every entry of the units vector goes on the trail
// (duplicates included) and its variable becomes inactive for later elections
__global__ void k_bcp_trail(const u32* __restrict__ units, DevCounters* dc, u32* __restrict__ trail, u32 trailCap,
                            unsigned char* __restrict__ vstate) {
    const u32 n = dc->numUnits, base = dc->trailSize;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 u = units[i];
        if (base + i < trailCap) trail[base + i] = u;
        vstate[LABS(u)] = 2;  // FROZEN_M (markFrozen, solver.hpp:515-519)
    }
}
__global__ void k_bcp_finish(DevCounters* dc, u32 trailCap) {
    const u32 n = dc->numUnits;
    dc->trailSize = min(dc->trailSize + n, trailCap);
    dc->numUnits = 0;   // (unassignedDec: counted where the variables are assigned, k_bcp_seed / k_bcp_level)
}
__global__ void k_bcp_reset(DevCounters* dc) { dc->bcpCurr = dc->bcpNext = dc->bcpConfl = dc->bcpLevel = 0; }

int runProp(Ctx* c, bool* conflict) {
    *conflict = false;
    if (!c->nUnits) return 0;
    u32* state = c->rank;  // reused: rank[] is rebuilt by the election that follows
    CUDA_TRY(cudaMemsetAsync(state, 0, (size_t)(c->V + 1) * 4, c->stream));
    LAUNCH(c, k_bcp_reset, 1, 1, 0, c->dc);
    LAUNCH(c, k_bcp_seed, gridFor(c->nUnits, 256), 256, 0, state, c->wlA, c->dc, c->eliminated, c->units, c->nUnits);
    const u32 unitsCap = 2 * (c->V + 1);
    for (int batch = 1;; batch = batch < 16 ? batch * 2 : 16) {
        for (int b = 0; b < batch; b++) {
            LAUNCH(c, k_bcp_level, 148 * 8, 256, 0, c->hdr[c->cur], c->pool[c->cur], c->otStart, c->otSize, c->occurs, state,
                   c->eliminated, c->wlA, c->wlB, c->dc, c->units, unitsCap);
            LAUNCH(c, k_bcp_advance, 1, 1, 0, c->dc);
        }
        int rc = syncCounters(c);
        if (rc) return rc;
        if (c->hdc->bcpConfl || !c->hdc->bcpCurr) break;
    }
    if (c->hdc->bcpConfl) { *conflict = true; return 0; }
    const u32 n = c->hdc->numCls;
    LAUNCH(c, k_bcp_apply, gridFor(n, 256), 256, 0, c->hdr[c->cur], c->pool[c->cur], n, state);
    const u32 nu = c->hdc->numUnits;
    LAUNCH(c, k_bcp_trail, gridFor(nu, 256), 256, 0, c->units, c->dc, c->trail, c->V + 1 + unitsCap, c->vstate);
    LAUNCH(c, k_bcp_finish, 1, 1, 0, c->dc, c->V + 1 + unitsCap);
    return 0;
}
