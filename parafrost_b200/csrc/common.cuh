// parafrost_b200/csrc/common.cuh -- shared definitions of the B200-native SIGmA engine.
//
// Data layout in HBM (one arena per context, carved once per sigma_load, see api.cu):
//   hdr[capC]   uint4 per clause  {x: word offset into pool, y: size, z: 32-bit signature,
//                                  w: the reference's SCLAUSE word 0 = st:2 f:1 a:1 u:2 lbd:26}
//               (replaces the interleaved SCLAUSE records + uint64 refs, src/gpu/sclause.cuh:37-42,
//                src/gpu/cnf.cuh:35-132; a clause *index* plays the role of S_REF: refs are
//                allocated in increasing order, so index order == ref order for every tie-break)
//   pool[capW]  uint32 literals, sorted ascending inside a clause
//   key[capC]   uint4 {size, first lit, last lit, sig}: the OLIST_CMP key (src/gpu/key.cuh:67-83),
//               rebuilt by the histogram pass so the list sort never chases clause pointers
//   hist[2V+2], otStart[2V+3], otSize[2V+2], occurs[capW]  occurrence table with 4-byte entries
//   otPairs[capW] uint2 {literal, clause}: the radix-partition buffer of the OT build (cnf.cu)
//               (replaces OT/OL with 8-byte refs + 16-byte list headers, src/gpu/table.cuh:32-79)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sigma.h"

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

// ---------------------------------------------------------------- literal helpers (constants.hpp:72-90)
#define LABS(l) ((l) >> 1)
#define LSIGN(l) ((l) & 1u)
#define LFLIP(l) ((l) ^ 1u)
#define V2L(v) ((v) << 1)
#define LNEG(l) ((l) | 1u)
#define MAPHASH(l) (1u << ((l) & 31u))
#define IS_TAUT(a, b) ((((a) ^ (b))) == 1u)
#define SUBSIG(a, b) (!((a) & ~(b)))

// ---------------------------------------------------------------- clause header word w (SCLAUSE word 0)
#define CB_LEARNT 1u
#define CB_DELETED 2u
#define CB_ST_MASK 3u
#define CB_MOLTEN 4u
#define CB_ADDED 8u
#define CB_USAGE_SHIFT 4
#define CB_USAGE_MASK (3u << 4)
#define CB_LBD_SHIFT 6
#define C_ORIGINAL(w) (((w) & CB_ST_MASK) == 0u)
#define C_LEARNT(w) ((w) & CB_LEARNT)
#define C_DELETED(w) ((w) & CB_DELETED)
#define C_MOLTEN(w) ((w) & CB_MOLTEN)

// eliminated[] byte (constants.cuh:33-36)
#define MELTING_MASK 1
#define ADDING_MASK 2
#define FORCED_MASK 4
// BVE type word (constants.cuh:37-62)
#define RES_MASK 1u
#define AOIX_MASK 2u
#define CORE_MASK 3u
#define ADDEDCLS_MAX 0x3FFFu
#define ADDEDLITS_MAX 0xFFFFu
#define ENCODEVARINFO(T, CLS, LITS) ((T) | ((CLS) << 2) | ((LITS) << 16))
#define RECOVERTYPE(x) ((x) & 3u)
#define RECOVERADDEDCLS(x) (((x) & 0xFFFCu) >> 2)
#define RECOVERADDEDLITS(x) (((x) & 0xFFFF0000u) >> 16)
#define NBUCKETS 3u        // header words of the reference's SCLAUSE (logical capacity arithmetic)
#define SUB_MAX_CL_SIZE 1000
#define LBD_TIER1 2
#define USAGET3 1u
#define MAXFUNVAR 12
#define FUNTABLEN 64

// MIS states (lcve.cu)
#define MIS_NONE 0
#define MIS_UNDECIDED 1
#define MIS_ELECTED 2
#define MIS_FROZEN 3
#define MIS_LIVESTOP 4
#define MIS_HALF 5      // not elected, but its positive-list neighbours are frozen: depFreeze_d succeeded on the positive list and hit a
                        // clause longer than lcveclausemax in the negative one, whose freezes alone are rolled back (lcve.cu:33-62, :96-98)
// candidate class
#define CS_NONE 0
#define CS_CAND 1
#define CS_STOP 2

#define NOVAR 0xFFFFFFFFu

// ---------------------------------------------------------------- device-resident scalars
struct DevCounters {
    u32 numCls;        // clause slots in use (== cnf->size())
    u32 poolUsed;      // words used in pool
    u64 dataSize;      // logical _data.size of the reference layout (words)
    u32 liveCls;       // count.cu
    u32 pad0;
    u64 liveLits;
    u32 numElected;
    u32 numUnits;      // vars->units size
    u32 resolvedSize;  // vars->resolved size
    u32 trailSize;
    int lastElimID;    // lastEliminatedID (cnf.cu:29)
    u32 misStopRank;
    u32 wlNext;        // MIS worklist append cursor
    u32 wlCnt[3];      // rotating MIS worklist sizes (lcve.cu)
    u32 firstStop;     // rank of the first bound-violating candidate (lcve.cu)
    u32 flags;         // bit0: resolved overflow, bit1: units overflow, bit2: hole after failed MEMORY_SAFE, bit3: a clause with >= 2^14 literals,
                       // bit4: a BVE candidate could trip the proof guard (resolve.cuh:66-70), bit5: proof stream overflow,
                       // bit6: resolvents fit the logical capacities but not the arena, bit7: bad literal / offsets in the input
    u32 bcpCurr, bcpNext, bcpConfl, bcpLevel;
    u32 nFrozen;       // number of first-frozen variables mapped into varcore (<= 12 needed)
    u32 unassignedDec; // variables assigned by prop()
    u32 sortCnt[9], sortCur[9];   // list-sort length classes (otsort.cu)
    u32 addedCls;      // resolvents appended by the last BVE
    u32 bin[4];        // group-size class sizes of the elected variables (elim.cu) + redo queue
    u32 scratch[16];   // 0,1 scan totals; 2 MIS push count; 3,4 ERE queue count / overflow; 5 load check; 6 max score; 7 big-bucket units; 8 MIS_HALF count; 9 freezer count
    u32 froz12[12];    // variables currently holding a function-table index in varcore
    // device DRAT stream (elim.cu, proof kernels): bytes appended this round, the reference's capacity, units mark
    u32 proofSize, proofCap, proofUnits0, proofPad;
    u64 proofLitBytes; // proof bytes of every literal of the loaded formula (cuPROOF::count, proof.cu:101-121)
    u64 profBytes;     // kernel profile mode: bytes walked by the per-variable kernels (sigma_kernel_stats)
    u64 orgCL[2];      // original clauses / their literals counted on the device (sigma_load_sclauses, sigma_continue)
};

struct KOpts {   // kernel-side options (replaces __constant__ kOpts, options.cuh:27-45)
    u32 ve_clause_max, xor_max_arity, sub_max_occurs, ere_max_occurs, bce_max_occurs, sh_max_bve_out1;
    int ere_clause_max;
    int ve_fun_en, ve_lbound_en, in_mode;
    int proof_en;      // opts.proof_en: SUB leaves the molten marks for the proof pass
    u32 refsCap;       // logical refs capacity
    u64 dataCap;       // logical data capacity (words)
    u32 physC;         // clause slots the arena really holds (hdr[]) ...
    u64 physW;         // ... and literal words (pool[]): the logical capacities may pass them after a GC (api.cu: buildOT)
};

#define KT_MAX_KERNELS 96
#define KT_POOL 2048
struct Ctx;
int  ktRegister(const char* name);       // api.cu: kernel name -> stable index
void ktBegin(Ctx* c, int id);
void ktEnd(Ctx* c);

// ---------------------------------------------------------------- context
struct Ctx {
    int device;
    cudaStream_t stream;
    sigma_opts o;
    char err[256];
    // arena
    char* arena; size_t arenaBytes, arenaUsed, arenaPeak; u64 cudaMallocs;
    // sizes
    u32 V, ND; u64 C0, L0; u32 capC; u64 capW; u32 resolvedCap;   // capC / capW: physical sizes of hdr[] / pool[]
    u64 logC, logW;    // the reference's logical capacities of awaken (simplify.cu:84-98) for the loaded formula and options
    u64 orgClauses, orgLiterals;
    bool loaded, begun, needReload;   // needReload: options set after sigma_load do not fit the carved arena (sigma_set_opts)
    // input (pristine)
    u32* inLits; u64* inOffs; u32* inMeta; u32* inMetaBuf; u64 inCapC, inCapL;   // inCap*: carved sizes (slack for sigma_continue)
    // CNF double buffer
    uint4* hdr[2]; u32* pool[2]; int cur;
    uint4* key;
    // OT
    u32 *hist, *otStart, *otSize, *occurs;
    uint2* otPairs; u32* otCur; u32* otBig; u32 otShift, otNB;
    // OT build v2 (cnf.cu): ranks of the counting pass (8 x 16 bits per clause), count matrix [tiles][NBp], bucket totals / starts
    uint4* rk8; u32* cntMat; u32* runMat; u32* otSeg; u32* bstart; u32 otNBp, otCPT, otTiles; u64 otTilePairs; bool otHot; bool attrOT2, attrTma;   // partition buffer of the OT build (cnf.cu): (literal, clause) pairs, bucket cursors
    // vars
    u32 *scores, *eligible, *rank, *sortK, *sortV, *elected, *units, *resolved, *trail, *vorg, *varcore;
    unsigned char *mis, *cstat, *vstate, *vstate0, *assumed, *assumedBuf, *eliminated, *needSort;
    u64 resolvedCapPhys;
    u32 *wlA, *wlB;
    // BVE arrays
    u32 *veType, *veUcnt, *veRpos, *veRes, *veUoff, *veResOff; u64* veRref;
    // scan / misc scratch
    u32 *scanTmp; u64* scanTmp64; u32 *flagA, *flagB; u64* flag64;
    u32 *radixHist; u32 *qMed;   // qMed: literals grouped by list-length class (otsort.cu)
    DevCounters* dc; DevCounters* hdc;   // device + pinned host mirror
    // host-side loop state (simplify.cu:136-241)
    int phase, multiplier, simpstate, cnfstate; bool compacted;
    u64 numClauses, numLiterals; u32 nUnits, numElected, currMelted; i64 unassigned;
    u64 refsCap, dataCap;
    i64 cdiff, ldiff, clsbefore, litsbefore;
    bool loopDone;
    sigma_round_report* rounds; u32 nRounds, capRounds;
    sigma_stage_reduction* reds; u32 nReds, capReds;   // opts.log_reductions (LOGREDALL / LOGREDCL)
    u64 launches;
    float stageMs[16];
    cudaEvent_t ev0, ev1, evRun0, evRun1; bool ownStream;
    double msTotal;
    u32 lastElectedCount;
    u32 lastPropSeeds, lastPropTrail0, lastPropTotal;   // the last prop(): BVE-origin units, trail size before it, entries it appended
    bool varcoreDead, attrSort, attrElim, attrOT;
    bool histFresh;    // key[] and the count matrix were produced by the awaken pass (k_ot_count) and the store is untouched since
    bool countsFresh;  // hdc->liveCls / liveLits describe the clause store as it is now
    bool otValid;      // the occurrence table built last round still describes the clause store (api.cu)
    i64 unassigned0;   // unassigned variables of the loaded formula (inf.unassigned)
    // per-kernel CUDA-event timing (sigma_kernel_profile): event pairs recorded on the launch stream
    bool ktOn; cudaEvent_t* ktEv; int* ktId; u32 ktUsed;
    // device DRAT stream: device buffer + deleted-flag snapshot (arena), pinned staging, host chunk store
    unsigned char* proofBuf; u32* proofSnap; u64 proofPhys; u32 proofBMax; bool proofCarved;
    unsigned char* proofHost; u64 proofHostCap;
    unsigned char* proofAll; u64 proofAllSize, proofAllCap; u64* proofEnds; u32 nProofChunks, capProofChunks;
    sigma_proof_sink proofSink; void* proofUser;
    float ktMs[KT_MAX_KERNELS]; u32 ktCount[KT_MAX_KERNELS];
    double ktBytes[KT_MAX_KERNELS]; int ktLastId;   // algorithmic bytes per kernel name, fed by KB() / profGather()
};

enum Stage { ST_VO = 0, ST_SIG, ST_IO, ST_GC, ST_COT, ST_SOT, ST_ROT, ST_VE, ST_SUB, ST_BCE, ST_ERE, ST_PROP, ST_LCVE, ST_CNT };

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            snprintf(c->err, sizeof c->err, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -(int)_e;                                                                 \
        }                                                                                    \
    } while (0)

#define LAUNCH(c, kern, grid, block, smem, ...)                     \
    do {                                                            \
        static int _kid = -1;                                       \
        if ((c)->ktOn) { if (_kid < 0) _kid = ktRegister(#kern); ktBegin((c), _kid); (c)->ktLastId = _kid; } \
        kern<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__); \
        if ((c)->ktOn) ktEnd(c);                                    \
        (c)->launches++;                                            \
    } while (0)

// algorithmic bytes of the launch just made (kernel profile mode only): what the kernel must move at least -
// 16-byte clause headers, 4-byte literals and list entries
#define KB(c, expr) do { if ((c)->ktOn && (c)->ktLastId >= 0) (c)->ktBytes[(c)->ktLastId] += (double)(expr); } while (0)

static inline u32 divup(u64 a, u32 b) { return (u32)((a + b - 1) / b); }
// grid-stride launches: enough CTAs to fill 148 SMs several times over, never more than the work
static inline u32 gridFor(u64 n, u32 block, u32 perThread = 1) {
    u64 g = (n + (u64)block * perThread - 1) / ((u64)block * perThread);
    const u64 cap = 148ull * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (u32)g;
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ u32 laneId() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemaskLt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// warp-aggregated atomic append; returns the slot of this lane (all active lanes must call)
__device__ __forceinline__ u32 warpAggInc(u32* counter) {
    const u32 mask = __activemask();
    const u32 leader = __ffs(mask) - 1;
    u32 base = 0;
    if (laneId() == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & lanemaskLt());
}

__device__ __forceinline__ u32 warpSum(u32 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u32 warpMax(u32 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// inclusive warp scan
__device__ __forceinline__ u32 warpIncl(u32 v) {
    const u32 l = laneId();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, v, o); if (l >= (u32)o) v += t; }
    return v;
}

// ---------------------------------------------------------------- launchers (one per stage)
// scan.cu
void scanExclusiveU32(Ctx* c, const u32* in, u32* out, u64 n, u32 init, u32* totalOut /*device, may be null*/);
void scanExclusiveU64(Ctx* c, const u64* in, u64* out, u64 n, u64 init);
// cnf.cu
void launchAwaken(Ctx* c);
void launchHistKey(Ctx* c);
void launchScatter(Ctx* c);
void launchCount(Ctx* c);
void launchGC(Ctx* c);
int  launchStore(Ctx* c, u64* nCls, u64* nLits, int form, bool writeBackOrder);   // form: 0 arrays (bits, sig, offs), 1 SCLAUSE records, 2 compact (bits, sizes);   // writeBackOrder: apply -aggresivesort (cacheCNF only)
// otsort.cu
void launchSortOT(Ctx* c, int mode);   // 0 all lists, 1 elected variables, 2 lists flagged in needSort
// lcve.cu
int  runLCVE(Ctx* c);
// stable LSD radix sort of (key, value) pairs on the low `bits` bits; result in (keys, vals); n <= max(V + 1, capC)
void radixSortPairs(Ctx* c, u32* keys, u32* vals, u32* keys2, u32* vals2, u32 n, u32 bits);
// prop.cu
int  runProp(Ctx* c, bool* conflict);
// elim.cu
void launchSUB(Ctx* c, const KOpts& k);
void launchVE(Ctx* c, const KOpts& k);
void launchBCE(Ctx* c, const KOpts& k);
void launchERE(Ctx* c, const KOpts& k);
void launchProofCount(Ctx* c);   // cuPROOF::count: dc->proofLitBytes, dc->proofCap
// api.cu
int  syncCounters(Ctx* c);   // D2H of DevCounters into c->hdc, stream synchronised
// elim.cu, profile mode: sum over the variables wl[0 .. *count) (indices into elected[], or variables when `direct`) of
// (4 + 16 + 4|c|) over the clauses of both occurrence lists -> *out (host), the SUB / BVE / ERE / MIS byte formula of SURVEY 8d
double profGather(Ctx* c, const u32* wl, const u32* countDev, u32 upper, bool direct);
