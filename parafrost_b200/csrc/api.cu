// parafrost_b200/csrc/api.cu -- C ABI (include/sigma.h), device arena and the round loop.
//
// Host driver = Solver::simplifying (src/gpu/simplify.cu:136-241) restated over the kernels of
// this directory.  Memory = one cudaMalloc per context, carved by a bump allocator when a
// formula is loaded (replaces cuMM + cuArena, src/gpu/memory.cu:99-387): no allocation, free or
// resize happens inside the round loop, and the arena is reused by later loads that fit.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

struct sigma_ctx : Ctx {};

static double nowMs() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ per-kernel event timing
#include <mutex>
static std::mutex gKtMutex;
static char gKtNames[KT_MAX_KERNELS][64];
static int gKtN = 0;
int ktRegister(const char* rawName) {
    // the LAUNCH macro stringifies its argument: "(k_ot_part2<3, 3>)" -> "k_ot_part2<3, 3>"
    char name[64];
    const size_t len = strlen(rawName);
    const bool paren = len >= 2 && rawName[0] == '(' && rawName[len - 1] == ')';
    snprintf(name, sizeof name, "%.*s", (int)(paren ? len - 2 : len), rawName + (paren ? 1 : 0));
    std::lock_guard<std::mutex> lk(gKtMutex);
    for (int i = 0; i < gKtN; i++) if (!strncmp(gKtNames[i], name, 63)) return i;
    if (gKtN == KT_MAX_KERNELS - 1) return KT_MAX_KERNELS - 1;   // overflow bucket
    strncpy(gKtNames[gKtN], name, 63);
    return gKtN++;
}
static void ktFlush(Ctx* c) {
    if (!c->ktUsed) return;
    cudaEventSynchronize(c->ktEv[2 * (c->ktUsed - 1) + 1]);
    for (u32 i = 0; i < c->ktUsed; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ktEv[2 * i], c->ktEv[2 * i + 1]) == cudaSuccess) { c->ktMs[c->ktId[i]] += ms; c->ktCount[c->ktId[i]]++; }
    }
    c->ktUsed = 0;
}
void ktBegin(Ctx* c, int id) {
    if (c->ktUsed == KT_POOL) ktFlush(c);
    c->ktId[c->ktUsed] = id;
    cudaEventRecord(c->ktEv[2 * c->ktUsed], c->stream);
}
void ktEnd(Ctx* c) { cudaEventRecord(c->ktEv[2 * c->ktUsed + 1], c->stream); c->ktUsed++; }

extern "C" int sigma_kernel_profile(sigma_ctx* c, int enable) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    if (enable && !c->ktEv) {
        c->ktEv = (cudaEvent_t*)calloc(2 * KT_POOL, sizeof(cudaEvent_t));
        c->ktId = (int*)calloc(KT_POOL, sizeof(int));
        if (!c->ktEv || !c->ktId) return SIGMA_AWAKEN_FAIL;
        for (int i = 0; i < 2 * KT_POOL; i++) CUDA_TRY(cudaEventCreate(&c->ktEv[i]));
    }
    if (c->ktOn) ktFlush(c);
    if (enable == 2 || !enable) { /* keep totals */ } else { memset(c->ktMs, 0, sizeof c->ktMs); memset(c->ktCount, 0, sizeof c->ktCount); memset(c->ktBytes, 0, sizeof c->ktBytes); }
    c->ktLastId = -1;
    c->ktOn = enable != 0;
    return SIGMA_OK;
}
extern "C" int sigma_kernel_stats(sigma_ctx* c, char* names, float* ms, uint32_t* counts, double* bytes, uint32_t* n);
extern "C" int sigma_kernel_times(sigma_ctx* c, char* names, float* ms, uint32_t* counts, uint32_t* n) {
    return sigma_kernel_stats(c, names, ms, counts, nullptr, n);
}
extern "C" int sigma_kernel_stats(sigma_ctx* c, char* names, float* ms, uint32_t* counts, double* bytes, uint32_t* n) {
    if (!c || !n) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->ktOn) ktFlush(c);
    u32 out = 0;
    const u32 cap = *n;
    std::lock_guard<std::mutex> lk(gKtMutex);
    for (int i = 0; i < KT_MAX_KERNELS && out < cap; i++) {
        if (!c->ktCount[i]) continue;
        if (names) { strncpy(names + 64 * out, i < gKtN ? gKtNames[i] : "(other)", 63); names[64 * out + 63] = 0; }
        if (ms) ms[out] = c->ktMs[i];
        if (counts) counts[out] = c->ktCount[i];
        if (bytes) bytes[out] = c->ktBytes[i];
        out++;
    }
    *n = out;
    return SIGMA_OK;
}

// ------------------------------------------------------------------ options
extern "C" void sigma_default_opts(sigma_opts* o) {
    memset(o, 0, sizeof *o);
    o->phases = 5; o->ve_en = 1; o->ve_plus_en = 1; o->sub_en = 1; o->bce_en = 0; o->ere_en = 1; o->all_en = 0;
    o->mu_pos = 32; o->mu_neg = 32; o->lcve_min_vars = 2; o->lcve_max_occurs = 3000; o->lcve_clause_max = 30000;
    o->phase_lits_min = 500; o->shrink_rate = 2; o->lits_mul = 1.0;
    o->ve_fun_en = 1; o->ve_lbound_en = 0; o->ve_clause_max = 100; o->xor_max_arity = 10;
    o->ere_clause_max = 250; o->ere_max_occurs = 3000; o->sub_max_occurs = 3000; o->bce_max_occurs = 3000;
    o->sh_max_bve_out1 = 250; o->sigma_calls = 1; o->final_gc = 1; o->profile = 0;
    o->lcve_fast = 0;   // the reference CLI's default is 1; parity needs the deterministic walk (sigma.h)
}
extern "C" void sigma_normalize_opts(sigma_opts* o) {  // options.cpp:291-296
    o->ve_en = o->ve_en || o->ve_plus_en;
    if (o->all_en) o->ve_en = 1, o->ve_plus_en = 1, o->bce_en = 1, o->ere_en = 1;
    if (!o->phases && (o->ve_en || o->sub_en || o->bce_en)) o->phases = 1;
    if (o->phases && !(o->ve_en || o->sub_en || o->bce_en)) o->phases = 0;
    if (o->phases > 1 && !o->ve_en) o->phases = 1;
    if (o->ere_clause_max > 250) o->ere_clause_max = 250;
    if (o->sh_max_bve_out1 > 250) o->sh_max_bve_out1 = 250;
}
extern "C" const char* sigma_version(void) { return "sigma-b200 0.1 (sm_100a)"; }
extern "C" const char* sigma_last_error(const sigma_ctx* c) { return c ? c->err : "null context"; }

// ------------------------------------------------------------------ context
extern "C" int sigma_create(int device, const sigma_opts* o, sigma_ctx** out) {
    if (!out) return SIGMA_BAD_ARGUMENT;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) return e != cudaSuccess ? -(int)e : SIGMA_BAD_ARGUMENT;
    sigma_ctx* c = new (std::nothrow) sigma_ctx();
    if (!c) return SIGMA_AWAKEN_FAIL;
    memset(static_cast<Ctx*>(c), 0, sizeof(Ctx));
    c->device = device;
    if (o) c->o = *o; else sigma_default_opts(&c->o);
    sigma_normalize_opts(&c->o);
    c->cnfstate = SIGMA_UNSOLVED;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMallocHost(&c->hdc, sizeof(DevCounters))) != cudaSuccess || (e = cudaEventCreate(&c->ev0)) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&c->evRun0)) != cudaSuccess || (e = cudaEventCreate(&c->evRun1)) != cudaSuccess) {
        delete c;
        return -(int)e;
    }
    memset(c->hdc, 0, sizeof(DevCounters));
    c->ownStream = true;
    *out = c;
    return SIGMA_OK;
}

static bool logicalCaps(const Ctx* c, u64 num_clauses, u64 L0, u64 orgC, u64 orgL, u64* logC, u64* logW);

extern "C" int sigma_set_opts(sigma_ctx* c, const sigma_opts* o) {
    if (!c || !o) return SIGMA_BAD_ARGUMENT;
    const sigma_opts old = c->o;
    c->o = *o;
    sigma_normalize_opts(&c->o);
    (void)old;
    if (c->loaded) {
        // the arena was carved for the options in force at sigma_load (prepareLoad): options that enlarge the logical
        // capacities (ve_en, lits_mul, phases) or need the proof buffer take effect with the next sigma_load; until then
        // sigma_begin refuses to run instead of overrunning the arena
        u64 lc = 0, lw = 0;
        const bool ok = logicalCaps(c, c->C0, c->L0, c->orgClauses, c->orgLiterals, &lc, &lw) && lc <= c->capC && lw <= c->capW &&
                        (!c->o.proof_en || c->proofCarved);
        c->needReload = !ok;
        if (ok) { c->logC = lc; c->logW = lw; }
    }
    return SIGMA_OK;
}

extern "C" int sigma_set_stream(sigma_ctx* c, void* cuda_stream) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->ownStream) { cudaStreamDestroy(c->stream); c->ownStream = false; }
    c->stream = (cudaStream_t)cuda_stream;
    return SIGMA_OK;
}

extern "C" int sigma_destroy(sigma_ctx* c) {
    if (!c) return SIGMA_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->arena) cudaFree(c->arena);
    if (c->hdc) cudaFreeHost(c->hdc);
    if (c->proofHost) cudaFreeHost(c->proofHost);
    free(c->proofAll); free(c->proofEnds);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evRun0) cudaEventDestroy(c->evRun0);
    if (c->evRun1) cudaEventDestroy(c->evRun1);
    if (c->stream && c->ownStream) cudaStreamDestroy(c->stream);
    if (c->ktEv) { for (int i = 0; i < 2 * KT_POOL; i++) if (c->ktEv[i]) cudaEventDestroy(c->ktEv[i]); free(c->ktEv); free(c->ktId); }
    free(c->rounds); free(c->reds);
    delete c;
    return SIGMA_OK;
}

int syncCounters(Ctx* c) {
    CUDA_TRY(cudaMemcpyAsync(c->hdc, c->dc, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------ arena
struct Carver {
    char* base; size_t off;
    template <typename T> T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

// lays every buffer out; with base == nullptr it only measures
static size_t carve(Ctx* c, char* base) {
    Carver a{base, 0};
    const size_t V1 = (size_t)c->V + 1, ND = c->ND, capC = c->capC, capW = c->capW;
    const size_t nflag = (capC > V1 ? capC : V1) + 2;
    c->inLits = a.take<u32>(c->inCapL + 1);
    c->inOffs = a.take<u64>(c->inCapC + 1);
    c->inMetaBuf = a.take<u32>(c->inCapC + 1);
    c->inMeta = c->inMetaBuf;
    for (int b = 0; b < 2; b++) { c->hdr[b] = a.take<uint4>(capC + 1); c->pool[b] = a.take<u32>(capW + 4); }
    c->key = a.take<uint4>(capC + 1);
    c->hist = a.take<u32>(ND + 2); c->otStart = a.take<u32>(ND + 2); c->otSize = a.take<u32>(ND + 2);
    c->occurs = a.take<u32>(capW + 4);
    c->otPairs = a.take<uint2>(capW + 4); c->otCur = a.take<u32>(8192 + 2);
    c->otBig = a.take<u32>(8192 + capW / 32768 + 64 + 8192 + 8);   // work units of oversized buckets (k_ot_place_big) + their bucket ids
    c->rk8 = a.take<uint4>(capC + 1);
    c->cntMat = a.take<u32>(((size_t)capC / (1024 * 3) + 2) * 8192);   // one row of <= 8192 bucket counts per tile of >= 3072 clauses
    c->runMat = a.take<u32>(((size_t)capC / (1024 * 3) + 2) * 8192);   // ... and of run starts (column scan of the counts)
    c->otSeg = a.take<u32>((size_t)64 * 8192 + 8); c->bstart = a.take<u32>(8192 + 8);
    c->scores = a.take<u32>(V1); c->eligible = a.take<u32>(V1); c->rank = a.take<u32>(V1);
    c->sortK = a.take<u32>(V1); c->sortV = a.take<u32>(V1); c->elected = a.take<u32>(V1);
    c->units = a.take<u32>(2 * V1); c->trail = a.take<u32>(3 * V1);
    c->resolved = a.take<u32>((size_t)c->resolvedCapPhys + 2);
    c->vorg = a.take<u32>(V1); c->varcore = a.take<u32>(V1);
    c->mis = a.take<unsigned char>(V1); c->cstat = a.take<unsigned char>(V1);
    c->vstate = a.take<unsigned char>(V1); c->vstate0 = a.take<unsigned char>(V1);
    c->assumedBuf = a.take<unsigned char>(V1); c->assumed = c->assumedBuf; c->eliminated = a.take<unsigned char>(V1);
    c->needSort = a.take<unsigned char>(ND + 4);
    c->wlA = a.take<u32>(V1); c->wlB = a.take<u32>(V1);
    c->veType = a.take<u32>(V1); c->veUcnt = a.take<u32>(V1); c->veRpos = a.take<u32>(V1); c->veRref = a.take<u64>(V1);
    const size_t maxScan = (nflag > ND + 2 ? nflag : ND + 2);
    const size_t radixBlocks = (V1 > capC ? V1 : capC) / 4096 + 2;   // the election sorts V scores, -aggresivesort up to capC clause keys
    const size_t scanTiles = (maxScan > 256 * radixBlocks ? maxScan : 256 * radixBlocks) / 2048 + 4;
    c->scanTmp = a.take<u32>(scanTiles); c->scanTmp64 = a.take<u64>(scanTiles);
    c->flagA = a.take<u32>(nflag); c->flagB = a.take<u32>(nflag); c->flag64 = a.take<u64>(capC + 2);
    c->radixHist = a.take<u32>(256 * radixBlocks);
    c->qMed = a.take<u32>(ND);
    if (c->proofCarved) { c->proofBuf = a.take<unsigned char>(c->proofPhys + 64); c->proofSnap = a.take<u32>(capC / 32 + 2); }
    else { c->proofBuf = nullptr; c->proofSnap = nullptr; }
    c->dc = a.take<DevCounters>(1);
    return a.off + 256;
}

__global__ void k_iota(u32* __restrict__ a, u32 n) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = i;
}

// ------------------------------------------------------------------ load
// the reference's logical capacities for a formula under the context's options (simplify.cu:84-98)
static bool logicalCaps(const Ctx* c, u64 num_clauses, u64 L0, u64 orgC, u64 orgL, u64* logC, u64* logW) {
    u64 numCls = num_clauses, numLits = L0;
    if (c->o.phases) {
        numCls += c->o.ve_en ? orgC : 0;
        numLits += c->o.ve_en ? (u64)((double)orgL * c->o.lits_mul) : 0;
    }
    if (numCls >= 0xFFFFFFF0ull || numCls * NBUCKETS + numLits >= 0xFFFFFFF0ull) return false;
    *logC = numCls; *logW = numCls * NBUCKETS + numLits;
    return true;
}
// sizes the logical capacities and the arena for a formula of num_clauses clauses / L0 literals
static int prepareLoad(Ctx* c, uint32_t max_var, uint64_t num_clauses, u64 L0, u64 orgC, u64 orgL, const uint32_t* vorg) {
    // device DRAT stream: the logical capacity is the reference's (1.5 x the proof bytes of the input literals, counted on
    // the device at sigma_begin); the buffer is sized here from the widest literal, which bounds it
    c->proofCarved = c->o.proof_en != 0;
    if (c->proofCarved) {
        u32 maxOrg = max_var;
        if (vorg) { maxOrg = 1; for (size_t v = 1; v <= max_var; v++) if (vorg[v] > maxOrg) maxOrg = vorg[v]; }
        if (maxOrg >= (1u << 30)) return SIGMA_BAD_ARGUMENT;
        u32 widest = 2 * maxOrg + 1, b = 1;
        while (widest & 0xFFFFFF80u) { b++; widest >>= 7; }
        c->proofBMax = b;
        c->proofPhys = (u64)(1.5 * (double)b * (double)L0) + 16;
        if (c->proofPhys >= 0xFFFFFF00ull) return SIGMA_AWAKEN_FAIL;   // the reference's proof capacity is a uint32 (simplify.cu:130)
    }
    // election words carry a 27-bit rank (lcve.cu); clause indices are 32-bit
    if (max_var >= (1u << 27) - 2 || num_clauses >= 0xFFFFFFF0ull) return SIGMA_BAD_ARGUMENT;
    c->V = max_var; c->ND = 2 * (max_var + 1);
    c->C0 = num_clauses; c->L0 = L0;
    c->inCapC = num_clauses + num_clauses / 8 + 1024; c->inCapL = L0 + L0 / 8 + 4096;   // input arrays: room for the clauses a sigma_continue appends
    c->needReload = false;
    c->orgClauses = orgC; c->orgLiterals = orgL;
    // logical capacities of awaken (simplify.cu:84-98)
    if (!logicalCaps(c, num_clauses, L0, orgC, orgL, &c->logC, &c->logW)) return SIGMA_CNFALLOC_FAIL;
    // Physical sizes: at least the logical ones.  reallocCNF(true) (cnf.cu:129-144) moves the logical capacities to
    // 2 x the live clauses (+ literals) at every GC; the reference reallocates, this arena does not, so it is sized
    // for that up front as far as the input tells (learnt clauses loaded: 2 C0 > C0 + originals).  A resolvent batch
    // that fits the logical capacities but not the arena fails the round loudly (flags bit 6), never silently.
    u64 numCls = c->logC;
    if (c->o.phases && c->o.ve_en && 2 * num_clauses > numCls) numCls = 2 * num_clauses;
    numCls += numCls / 8 + 1024;                // slack: the clauses a sigma_continue appends, capacities that grow a little after a GC
    const u64 numWords = c->logW + (numCls - c->logC) * NBUCKETS + c->logW / 8;
    if (numCls >= 0xFFFFFFF0ull || numWords >= 0xFFFFFFF0ull) return SIGMA_CNFALLOC_FAIL;
    c->capC = (u32)numCls;
    c->capW = numWords;                         // data cap in words: a pool this big can never overflow below the logical caps
    const u64 rc = num_clauses + L0;            // savedLits (simplify.cu:85)
    c->resolvedCap = (u32)(rc > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : rc);
    { const u64 rp = rc + rc / 8 + 4096; c->resolvedCapPhys = rp > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : rp; }   // slack for sigma_continue
    const size_t need = carve(c, nullptr);
    if (need > c->arenaBytes) {
        if (c->arena) { CUDA_TRY(cudaFree(c->arena)); c->arena = nullptr; c->arenaBytes = 0; }
        cudaError_t e = cudaMalloc(&c->arena, need);
        if (e != cudaSuccess) {
            cudaGetLastError();
            snprintf(c->err, sizeof c->err, "arena of %zu bytes: %s", need, cudaGetErrorString(e));
            return SIGMA_CNFALLOC_FAIL;
        }
        c->arenaBytes = need;
        c->cudaMallocs++;
    }
    c->arenaUsed = carve(c, c->arena);
    if (c->arenaUsed > c->arenaPeak) c->arenaPeak = c->arenaUsed;
    return SIGMA_OK;
}
// per-variable inputs + completion of a load
static int finishLoad(Ctx* c, const uint32_t* vorg, const uint8_t* vstate, const uint8_t* assumed) {
    const size_t V1 = (size_t)c->V + 1;
    if (vorg) CUDA_TRY(cudaMemcpyAsync(c->vorg, vorg, V1 * 4, cudaMemcpyHostToDevice, c->stream));
    else LAUNCH(c, k_iota, gridFor(V1, 256), 256, 0, c->vorg, (u32)V1);
    if (vstate) CUDA_TRY(cudaMemcpyAsync(c->vstate0, vstate, V1, cudaMemcpyHostToDevice, c->stream));
    else CUDA_TRY(cudaMemsetAsync(c->vstate0, 0, V1, c->stream));
    if (assumed) CUDA_TRY(cudaMemcpyAsync(c->assumed, assumed, V1, cudaMemcpyHostToDevice, c->stream));
    else c->assumed = nullptr;
    if (c->proofCarved && c->proofHostCap < c->proofPhys) {   // pinned mirror of the device stream (cuPROOF::alloc, proof.cu:201-230)
        if (c->proofHost) { cudaFreeHost(c->proofHost); c->proofHost = nullptr; c->proofHostCap = 0; }
        CUDA_TRY(cudaMallocHost(&c->proofHost, c->proofPhys));
        c->proofHostCap = c->proofPhys;
    }
    i64 un = c->V;
    if (vstate) for (size_t v = 1; v < V1; v++) if (vstate[v]) un--;
    c->unassigned0 = un;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->loaded = true; c->begun = false;
    return SIGMA_OK;
}

extern "C" int sigma_load(sigma_ctx* c, uint32_t max_var, uint64_t num_clauses, const uint32_t* lits,
                          const uint64_t* offs, const uint32_t* meta, const uint32_t* vorg, const uint8_t* vstate,
                          const uint8_t* assumed) {
    if (!c || !lits || !offs || !max_var) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    const u64 L0 = offs[num_clauses];
    // stats.clauses.original / stats.literals.original (solver.hpp:164-165)
    u64 orgC = num_clauses, orgL = L0;
    if (meta) {
        orgC = 0; orgL = 0;
        for (u64 i = 0; i < num_clauses; i++) if (!(meta[i] & CB_LEARNT)) { orgC++; orgL += offs[i + 1] - offs[i]; }
    }
    int rc = prepareLoad(c, max_var, num_clauses, L0, orgC, orgL, vorg);
    if (rc) return rc;
    // host -> device (extractCNF + reflectCNF, cnf.cu:166-184)
    CUDA_TRY(cudaMemcpyAsync(c->inLits, lits, L0 * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->inOffs, offs, (num_clauses + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (meta) CUDA_TRY(cudaMemcpyAsync(c->inMeta, meta, num_clauses * 4, cudaMemcpyHostToDevice, c->stream));
    else c->inMeta = nullptr;
    return finishLoad(c, vorg, vstate, assumed);
}

// 32-bit offsets over PCIe, widened on the device (staged in rk8[], free until the first counting pass)
__global__ void k_widen_offs(const u32* __restrict__ in, u64 n, u64* __restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}
extern "C" int sigma_load32(sigma_ctx* c, uint32_t max_var, uint64_t num_clauses, const uint32_t* lits,
                            const uint32_t* offs32, const uint32_t* meta, const uint32_t* vorg, const uint8_t* vstate,
                            const uint8_t* assumed) {
    if (!c || !lits || !offs32 || !max_var) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    const u64 L0 = offs32[num_clauses];
    u64 orgC = num_clauses, orgL = L0;
    if (meta) {
        orgC = 0; orgL = 0;
        for (u64 i = 0; i < num_clauses; i++) if (!(meta[i] & CB_LEARNT)) { orgC++; orgL += offs32[i + 1] - offs32[i]; }
    }
    int rc = prepareLoad(c, max_var, num_clauses, L0, orgC, orgL, vorg);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->inLits, lits, L0 * 4, cudaMemcpyHostToDevice, c->stream));
    u32* stage = (u32*)c->rk8;   // 16 bytes per clause slot
    CUDA_TRY(cudaMemcpyAsync(stage, offs32, (num_clauses + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_widen_offs, gridFor(num_clauses + 1, 256, 4), 256, 0, stage, num_clauses + 1, c->inOffs);
    if (meta) CUDA_TRY(cudaMemcpyAsync(c->inMeta, meta, num_clauses * 4, cudaMemcpyHostToDevice, c->stream));
    else c->inMeta = nullptr;
    return finishLoad(c, vorg, vstate, assumed);
}

// SCLAUSE records {word 0, sig, size, literals...} at refs[i] -> the engine's input arrays
__global__ void k_unpack_sclauses(const u32* __restrict__ data, const u64* __restrict__ refs, u64 C, u64 numWords,
                                  u32* __restrict__ inLits, u64* __restrict__ inOffs, u32* __restrict__ inMeta, u32* bad,
                                  unsigned long long* orgCL) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < C; i += (u64)gridDim.x * blockDim.x) {
        const u64 r = refs[i];
        // records are appended in ref order without gaps (cnf.cuh:82-97): literal offset = ref - 3 i
        if (r < NBUCKETS * i || r + NBUCKETS > numWords) { atomicOr(bad, 1u); continue; }
        const u32 sz = data[r + 2];
        if (data[r] & CB_DELETED) { atomicOr(bad, 1u); continue; }   // extractCNF never mirrors deleted clauses (cnf.cu:176-184)
        const u64 next = i + 1 < C ? refs[i + 1] : numWords;
        if (r + NBUCKETS + sz != next) { atomicOr(bad, 1u); continue; }
        const u64 o = r - NBUCKETS * i;
        inOffs[i] = o;
        if (i + 1 == C) inOffs[C] = o + sz;
        inMeta[i] = data[r] & ~(CB_DELETED | CB_MOLTEN | CB_ADDED);
        for (u32 k = 0; k < sz; k++) inLits[o + k] = data[r + NBUCKETS + k];
        if ((data[r] & CB_ST_MASK) == 0) { atomicAdd(&orgCL[0], 1ull); atomicAdd(&orgCL[1], (unsigned long long)sz); }
    }
}

extern "C" int sigma_load_sclauses(sigma_ctx* c, uint32_t max_var, uint64_t num_clauses, const uint32_t* data_words,
                                   uint64_t num_words, const uint64_t* refs, const uint32_t* vorg, const uint8_t* vstate,
                                   const uint8_t* assumed) {
    if (!c || !data_words || !refs || !max_var || num_words < NBUCKETS * num_clauses) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    const u64 L0 = num_words - NBUCKETS * num_clauses;
    // originals (stats.clauses.original / literals.original) are counted by the unpack kernel; the arena is sized for the
    // upper bound "every clause is original" first, the exact logical capacities follow below
    int rc = prepareLoad(c, max_var, num_clauses, L0, num_clauses, L0, vorg);
    if (rc) return rc;
    // the reference's own two copies (reflectCNF, cnf.cu:166-174): record stream and refs, staged in
    // the inactive clause buffer, then unpacked on the device
    u32* dData = c->pool[1];
    u64* dRefs = c->flag64;
    u32* bad = &c->dc->scratch[5];
    CUDA_TRY(cudaMemcpyAsync(dData, data_words, num_words * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dRefs, refs, num_clauses * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(bad, 0, 4, c->stream));
    unsigned long long* orgCL = (unsigned long long*)c->dc->orgCL;
    CUDA_TRY(cudaMemsetAsync(orgCL, 0, 16, c->stream));
    if (num_clauses)
        LAUNCH(c, k_unpack_sclauses, gridFor(num_clauses, 256), 256, 0, dData, dRefs, num_clauses, num_words, c->inLits, c->inOffs, c->inMeta, bad, orgCL);
    else CUDA_TRY(cudaMemsetAsync(c->inOffs, 0, 8, c->stream));
    u32 hbad = 0;
    unsigned long long horg[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(horg, orgCL, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->orgClauses = horg[0]; c->orgLiterals = horg[1];
    if (!logicalCaps(c, num_clauses, L0, c->orgClauses, c->orgLiterals, &c->logC, &c->logW)) return SIGMA_CNFALLOC_FAIL;
    if (hbad) { snprintf(c->err, sizeof c->err, "SCLAUSE stream is not a gap-free sequence of records in ref order"); return SIGMA_BAD_ARGUMENT; }
    return finishLoad(c, vorg, vstate, assumed);
}

// ------------------------------------------------------------------ round loop
static KOpts makeK(Ctx* c) {
    KOpts k;
    k.ve_clause_max = c->o.ve_clause_max; k.xor_max_arity = c->o.xor_max_arity;
    k.sub_max_occurs = c->o.sub_max_occurs; k.ere_max_occurs = c->o.ere_max_occurs; k.bce_max_occurs = c->o.bce_max_occurs;
    k.sh_max_bve_out1 = c->o.sh_max_bve_out1; k.ere_clause_max = c->o.ere_clause_max;
    k.ve_fun_en = c->o.ve_fun_en && !c->varcoreDead; k.ve_lbound_en = c->o.ve_lbound_en; k.in_mode = c->o.sigma_calls > 1;
    k.refsCap = (u32)(c->refsCap > 0xFFFFFFFFull ? 0xFFFFFFFFull : c->refsCap); k.dataCap = c->dataCap;
    k.physC = c->capC; k.physW = c->capW;
    k.proof_en = c->o.proof_en && c->proofCarved;
    return k;
}

struct StageTimer {
    Ctx* c; int st; bool on;
    StageTimer(Ctx* c_, int st_) : c(c_), st(st_), on(c_->o.profile != 0) { if (on) cudaEventRecord(c->ev0, c->stream); }
    ~StageTimer() {
        if (!on) return;
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->stageMs[st] += ms;
    }
};

extern "C" int sigma_begin(sigma_ctx* c) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    if (!c->loaded) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->needReload) {
        snprintf(c->err, sizeof c->err, "the options set after sigma_load need a larger arena (or the proof buffer: proof_en must be set before "
                 "sigma_load, the stream buffer is carved with the arena): load the formula again");
        return SIGMA_BAD_ARGUMENT;
    }
    const size_t V1 = (size_t)c->V + 1;
    c->cur = 0;
    c->phase = c->multiplier = 0; c->simpstate = SIGMA_OK; c->cnfstate = SIGMA_UNSOLVED; c->compacted = false;
    c->nUnits = 0; c->numElected = 0; c->currMelted = 0; c->varcoreDead = false;
    c->lastPropSeeds = c->lastPropTrail0 = c->lastPropTotal = 0;
    c->nReds = 0;
    c->nRounds = 0; c->loopDone = false; c->launches = 0; c->msTotal = 0; c->otValid = false; c->countsFresh = false; c->histFresh = false;
    memset(c->stageMs, 0, sizeof c->stageMs);
    c->unassigned = c->unassigned0;
    memset(c->hdc, 0, sizeof(DevCounters));
    c->hdc->numCls = (u32)c->C0; c->hdc->poolUsed = (u32)c->L0; c->hdc->dataSize = c->C0 * NBUCKETS + c->L0;
    c->hdc->lastElimID = -1; c->hdc->misStopRank = NOVAR;
    for (int i = 0; i < 12; i++) c->hdc->froz12[i] = NOVAR;
    CUDA_TRY(cudaMemcpyAsync(c->dc, c->hdc, sizeof(DevCounters), cudaMemcpyHostToDevice, c->stream));
    c->proofAllSize = 0; c->nProofChunks = 0;
    if (c->o.proof_en) {
        if (!c->proofCarved) { snprintf(c->err, sizeof c->err, "proof_en must be set before sigma_load (the stream buffer is carved with the arena)"); return SIGMA_BAD_ARGUMENT; }
        launchProofCount(c);   // cuPROOF::count + the 1.5 x capacity (simplify.cu:128-132)
    }
    CUDA_TRY(cudaMemcpyAsync(c->vstate, c->vstate0, V1, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->eliminated, 0, V1, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->varcore, 0xFF, V1 * 4, c->stream));
    // logical capacities (simplify.cu:84-98)
    c->refsCap = c->logC; c->dataCap = c->logW;
    c->numClauses = c->C0; c->numLiterals = c->L0;
    c->cdiff = INT64_MAX; c->ldiff = INT64_MAX;
    c->clsbefore = (i64)c->numClauses; c->litsbefore = (i64)c->numLiterals;
    { StageTimer t(c, ST_SIG); launchAwaken(c); }
    c->begun = true;
    if (!c->C0) c->loopDone = true;
    // alldisabled (solver.hpp:722)
    if (!c->o.phases && !(c->o.all_en | c->o.ere_en)) c->loopDone = true;
    return SIGMA_OK;
}

// cuPROOF::cacheProof + writeProof (proof.cu:160-199, 232-247): the round's device stream -> pinned host -> sink + chunk
// store; c->hdc must be fresh (syncCounters).  Also the place where the proof-mode failures surface.
static int flushProof(Ctx* c) {
    if (!c->o.proof_en) return SIGMA_OK;
    const u32 n = c->hdc->proofSize;
    if ((c->hdc->flags & 32u) || n > c->hdc->proofCap) {
        snprintf(c->err, sizeof c->err, "proof stream overflow: %u bytes in one round, capacity %u (the reference only asserts, vector.cu)", n, c->hdc->proofCap);
        return SIGMA_OVERFLOW;
    }
    if (c->hdc->flags & 16u) {
        snprintf(c->err, sizeof c->err, "a BVE candidate is large enough to trip the proof guard (ADDEDPROOF_MAX, resolve.cuh:66-70): not supported with proof_en");
        return SIGMA_OVERFLOW;
    }
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(c->proofHost, c->proofBuf, n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemsetAsync(&c->dc->proofSize, 0, 4, c->stream));   // header.clear() (proof.cu:239-240)
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->hdc->proofSize = 0;
        if (c->proofAllSize + n > c->proofAllCap) {
            u64 cap = c->proofAllCap ? c->proofAllCap * 2 : (1u << 16);
            while (cap < c->proofAllSize + n) cap *= 2;
            unsigned char* p = (unsigned char*)realloc(c->proofAll, cap);
            if (!p) return SIGMA_AWAKEN_FAIL;
            c->proofAll = p; c->proofAllCap = cap;
        }
        memcpy(c->proofAll + c->proofAllSize, c->proofHost, n);
        c->proofAllSize += n;
    }
    if (c->nProofChunks == c->capProofChunks) {
        const u32 cap = c->capProofChunks ? c->capProofChunks * 2 : 16;
        u64* p = (u64*)realloc(c->proofEnds, cap * sizeof(u64));
        if (!p) return SIGMA_AWAKEN_FAIL;
        c->proofEnds = p; c->capProofChunks = cap;
    }
    c->proofEnds[c->nProofChunks++] = c->proofAllSize;
    if (n && c->proofSink) c->proofSink(c->proofUser, c->proofHost, n);
    return SIGMA_OK;
}

// LOGREDALL / LOGREDCL (logging.hpp:152-158, count.cu:185-208): live counts after a stage, only with opts.log_reductions
static int logReduction(Ctx* c, u32 stage, u32 varsRemoved) {
    if (!c->o.log_reductions) return 0;
    launchCount(c);
    const int rc = syncCounters(c);
    if (rc) return rc;
    if (c->nReds == c->capReds) {
        c->capReds = c->capReds ? c->capReds * 2 : 32;
        c->reds = (sigma_stage_reduction*)realloc(c->reds, c->capReds * sizeof(sigma_stage_reduction));
    }
    sigma_stage_reduction& e = c->reds[c->nReds++];
    e.round = (u32)c->phase; e.stage = stage; e.vars_removed = varsRemoved; e.pad = 0;
    e.clauses_before = c->numClauses; e.literals_before = c->numLiterals;
    e.clauses = c->hdc->liveCls; e.literals = c->hdc->liveLits;
    return 0;
}
extern "C" int sigma_reduction_log(const sigma_ctx* c, sigma_stage_reduction* out, uint32_t* n) {
    if (!c || !n) return SIGMA_BAD_ARGUMENT;
    const u32 m = c->nReds < *n ? c->nReds : *n;
    if (out && m) memcpy(out, c->reds, m * sizeof(sigma_stage_reduction));
    *n = c->nReds;
    return SIGMA_OK;
}

static void pushRound(Ctx* c, const sigma_round_report& r) {
    if (c->nRounds == c->capRounds) {
        c->capRounds = c->capRounds ? c->capRounds * 2 : 16;
        c->rounds = (sigma_round_report*)realloc(c->rounds, c->capRounds * sizeof(sigma_round_report));
    }
    c->rounds[c->nRounds++] = r;
}

// (GC) -> histogram -> scan -> partition/place : reallocOT + reallocCNF + createOTAsync (simplify.cu:164-167).
// The reference counts before it compacts; compaction drops deleted clauses only, so the histogram
// is the same either way and compacting first saves one pass over the clause store.
static void buildOT(Ctx* c, bool withGC, bool* didGC) {
    if (withGC) {
        StageTimer t(c, ST_GC);
        // reallocCNF(true), cnf.cu:129-144: new logical capacities, then compact
        const u64 maxAddedCls = c->o.ve_en ? c->numClauses : 0;
        const u64 maxAddedLits = c->o.ve_en ? (u64)((double)c->orgLiterals * c->o.lits_mul) : 0;
        c->refsCap = c->numClauses + maxAddedCls;
        c->dataCap = c->refsCap * NBUCKETS + (c->numLiterals + maxAddedLits);
        launchGC(c);
        c->hdc->numCls = (u32)c->numClauses; c->hdc->poolUsed = (u32)c->numLiterals;
        c->hdc->dataSize = c->numClauses * NBUCKETS + c->numLiterals;
        c->compacted = true;
        if (didGC) *didGC = true;
    }
    if (!c->histFresh || withGC) { StageTimer t(c, ST_VO); launchHistKey(c); }
    c->histFresh = false;
    { StageTimer t(c, ST_COT); launchScatter(c); }
}

extern "C" int sigma_round(sigma_ctx* c, sigma_round_report* rep, int* done) {
    if (!c || !done) return SIGMA_BAD_ARGUMENT;
    if (!c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    sigma_round_report r;
    memset(&r, 0, sizeof r);
    r.round = (u32)c->phase;
    r.kind = 2;
    r.literals_in = c->numLiterals;
    *done = 1;
    if (c->loopDone || !(c->numClauses && c->numLiterals && !c->simpstate)) { c->loopDone = true; if (rep) *rep = r; return SIGMA_OK; }
    const double t0 = nowMs();
    int rc;
    bool didGC = false;
    c->countsFresh = false;
    // cnf.cu:146-150
    const int times = c->phase + 1;
    const bool gc = times > 1 && times != c->o.phases && c->o.shrink_rate > 0 && (times % c->o.shrink_rate) == 0;
    if (!gc) c->compacted = false;
    // A round that changed nothing (no elimination, resolvent, unit, deletion or strengthening) leaves
    // the clause store, and with it the occurrence table, exactly as this round would rebuild them -
    // the usual way into the last round (stop(): !cdiff && !ldiff).  Compacting a store without
    // deleted slots or shrunken clauses is the identity, so a due GC does not invalidate the table.
    const bool gcIdentity = c->hdc->numCls == c->numClauses && c->hdc->poolUsed == c->numLiterals &&
                            c->hdc->dataSize == c->numClauses * NBUCKETS + c->numLiterals;
    if (c->otValid && !c->nUnits && (!gc || gcIdentity)) {
        if (gc) {   // reallocCNF(true), cnf.cu:129-144: the logical capacities move as if it had compacted
            const u64 maxAddedCls = c->o.ve_en ? c->numClauses : 0;
            const u64 maxAddedLits = c->o.ve_en ? (u64)((double)c->orgLiterals * c->o.lits_mul) : 0;
            c->refsCap = c->numClauses + maxAddedCls;
            c->dataCap = c->refsCap * NBUCKETS + (c->numLiterals + maxAddedLits);
            c->compacted = true; didGC = true;
        }
    } else
        buildOT(c, gc, &didGC);
    c->otValid = false;
    r.gc = didGC;
    // prop (elimbcp.cu:144-215)
    if (c->nUnits) {
        StageTimer t(c, ST_PROP);
        bool conflict = false;
        r.propagated = c->nUnits;
        c->lastPropSeeds = c->nUnits; c->lastPropTrail0 = c->hdc->trailSize; c->lastPropTotal = 0;
        if ((rc = runProp(c, &conflict))) return rc;
        if (conflict) {
            c->cnfstate = SIGMA_UNSAT; c->loopDone = true;
            r.ms = (float)(nowMs() - t0); pushRound(c, r); if (rep) *rep = r;
            return SIGMA_OK;
        }
        launchCount(c);
        if ((rc = syncCounters(c))) return rc;
        if (c->o.log_reductions) { const u32 forced = c->hdc->unassignedDec; if ((rc = logReduction(c, 0, forced))) return rc; }   // "BCP Reductions", elimbcp.cu:203
        c->numClauses = c->hdc->liveCls; c->numLiterals = c->hdc->liveLits;
        c->unassigned -= (i64)c->hdc->unassignedDec;
        c->lastPropTotal = c->hdc->trailSize - c->lastPropTrail0;
        r.trail_added = c->lastPropTotal;
        CUDA_TRY(cudaMemsetAsync(&c->dc->unassignedDec, 0, 4, c->stream));
        c->nUnits = 0;
        if (c->numLiterals) buildOT(c, false, nullptr);
        else launchScatter(c);
    }
    if (!c->numClauses) { c->loopDone = true; r.ms = (float)(nowMs() - t0); r.clauses = 0; r.literals = 0; pushRound(c, r); if (rep) *rep = r; return SIGMA_OK; }
    // LCVE (lcve.cu:302-398)
    {
        StageTimer t(c, ST_LCVE);
        if ((rc = runLCVE(c))) return rc;
    }
    if (c->hdc->flags & 128u) {
        snprintf(c->err, sizeof c->err, "the loaded formula holds a literal outside [2, 2 max_var + 2) or non-monotone clause offsets");
        return SIGMA_BAD_ARGUMENT;
    }
    c->numElected = c->hdc->numElected;
    r.elected = c->numElected;
    c->lastElectedCount = c->numElected;
    if (c->o.ve_fun_en && !c->varcoreDead && c->hdc->nFrozen == 0) c->varcoreDead = true;  // mapFrozen, lcve.cu:405
    if (c->numElected < c->o.lcve_min_vars) {
        c->loopDone = true; r.ms = (float)(nowMs() - t0); r.clauses = c->numClauses; r.literals = c->numLiterals;
        pushRound(c, r); if (rep) *rep = r;
        return SIGMA_OK;
    }
    const KOpts k = makeK(c);
    // stop() solver.hpp:748-753
    const bool stop = (c->phase == c->o.phases) || (c->simpstate == SIGMA_CNFALLOC_FAIL) || (!c->cdiff && !c->ldiff) ||
                      (c->phase > 2 && c->ldiff <= c->o.phase_lits_min);
    // sortOT (segsort.cu:37-48).  SUB/BVE/BCE only walk the lists of the elected variables in order;
    // the last round's ERE sorts the lists it is going to search by itself (launchERE).
    if (!stop) { StageTimer t(c, ST_SOT); launchSortOT(c, 1); }
    if (stop) {
        r.kind = 1;
        const bool ereRan = c->o.ere_en && c->numElected;
        if (ereRan) { StageTimer t(c, ST_ERE); launchERE(c, k); }
        if (ereRan && (rc = logReduction(c, 4, 0))) return rc;
        c->loopDone = true;
        launchCount(c);
        if ((rc = syncCounters(c))) return rc;
        c->countsFresh = true;
        if (ereRan && (rc = flushProof(c))) return rc;   // elimination.cu:305-306
        r.clauses = c->hdc->liveCls; r.literals = c->hdc->liveLits;
        r.ms = (float)(nowMs() - t0);
        pushRound(c, r); if (rep) *rep = r;
        return SIGMA_OK;
    }
    r.kind = 0;
    if (c->o.sub_en || c->o.ve_plus_en) { StageTimer t(c, ST_SUB); launchSUB(c, k); }
    if ((c->o.sub_en || c->o.ve_plus_en) && (rc = logReduction(c, 1, 0))) return rc;
    if (c->o.ve_en) { StageTimer t(c, ST_VE); launchVE(c, k); }
    if (c->o.ve_en && c->o.log_reductions) {
        if ((rc = syncCounters(c))) return rc;
        if ((rc = logReduction(c, 2, c->lastElectedCount - c->hdc->numElected))) return rc;
    }
    if (c->o.bce_en) {
        // BCE runs over the surviving elected variables (elimination.cu:280-291)
        if (c->o.ve_en) { if ((rc = syncCounters(c))) return rc; c->numElected = c->hdc->numElected; }
        if (c->numElected) { StageTimer t(c, ST_BCE); launchBCE(c, k); }
        if (c->numElected && (rc = logReduction(c, 3, 0))) return rc;
    }
    { StageTimer t(c, ST_CNT); launchCount(c); }
    if ((rc = syncCounters(c))) return rc;
    c->countsFresh = true;
    if (c->hdc->flags & 3u) { snprintf(c->err, sizeof c->err, "device vector overflow (flags %u)", c->hdc->flags); return SIGMA_OVERFLOW; }
    if (c->hdc->flags & 64u) {
        snprintf(c->err, sizeof c->err, "resolvents fit the reference's logical capacities (%llu refs / %llu words after a GC) but not the arena "
                 "(%u / %llu): the reference would have reallocated here", (unsigned long long)c->refsCap, (unsigned long long)c->dataCap, c->capC,
                 (unsigned long long)c->capW);
        return SIGMA_OVERFLOW;
    }
    if ((rc = flushProof(c))) return rc;   // cacheProof / writeProof, simplify.cu:174-184
    // updateNumPVs (simplify.cu:35-41)
    const u32 remained = c->o.ve_en ? c->hdc->numElected : c->numElected;
    r.eliminated = c->lastElectedCount - remained;
    c->currMelted += r.eliminated;
    c->numElected = remained;
    r.resolvents = c->o.ve_en ? c->hdc->addedCls : 0;
    c->numClauses = c->hdc->liveCls; c->numLiterals = c->hdc->liveLits;
    c->cdiff = c->clsbefore - (i64)c->numClauses; c->clsbefore = (i64)c->numClauses;
    c->ldiff = c->litsbefore - (i64)c->numLiterals; c->litsbefore = (i64)c->numLiterals;
    c->nUnits = c->hdc->numUnits;
    r.units = c->nUnits;
    c->otValid = !r.eliminated && !r.resolvents && !c->nUnits && !c->cdiff && !c->ldiff;   // structure untouched: table reusable
    c->phase++; c->multiplier++;
    c->multiplier += (c->phase == c->o.phases);
    r.clauses = c->numClauses; r.literals = c->numLiterals;
    r.ms = (float)(nowMs() - t0);
    pushRound(c, r);
    if (rep) *rep = r;
    *done = 0;
    return SIGMA_OK;
}

extern "C" int sigma_finish(sigma_ctx* c, sigma_report* rep) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    if (!c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc;
    if (!c->countsFresh) {   // the last round ended with a count and nothing touched the clause store since
        launchCount(c);
        if ((rc = syncCounters(c))) return rc;
        c->countsFresh = true;
    }
    c->numClauses = c->hdc->liveCls; c->numLiterals = c->hdc->liveLits;
    // simplify.cu:198-209
    if (c->cnfstate == SIGMA_UNSOLVED && (c->unassigned <= 0 || !c->numClauses)) c->cnfstate = SIGMA_SAT;
    if (rep) {
        memset(rep, 0, sizeof *rep);
        rep->cnfstate = c->cnfstate; rep->simpstate = c->simpstate; rep->rounds = c->nRounds;
        rep->eliminated_vars = c->currMelted;
        rep->clauses = c->numClauses; rep->literals = c->numLiterals;
        rep->clauses_in = c->C0; rep->literals_in = c->L0;
        rep->resolved_words = c->hdc->resolvedSize; rep->trail_units = c->hdc->trailSize;
        rep->ms_total = c->msTotal;
        for (int i = 0; i < 16; i++) rep->stage_ms[i] = c->stageMs[i];
        rep->kernel_launches = c->launches;
    }
    return SIGMA_OK;
}

extern "C" int sigma_run(sigma_ctx* c, sigma_report* rep) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    const double t0 = nowMs();
    cudaSetDevice(c->device);
    cudaEventRecord(c->evRun0, c->stream);
    int rc = sigma_begin(c);
    if (rc) return rc;
    int done = 0;
    while (!done) { if ((rc = sigma_round(c, nullptr, &done))) return rc; }
    if ((rc = sigma_finish(c, rep))) return rc;
    cudaEventRecord(c->evRun1, c->stream);
    rc = cudaEventSynchronize(c->evRun1) == cudaSuccess ? 0 : -1;
    c->msTotal = nowMs() - t0;
    if (rc) return rc;
    float dms = 0;
    cudaEventElapsedTime(&dms, c->evRun0, c->evRun1);
    if (rep) { rep->ms_total = c->msTotal; rep->ms_device = dms; }
    return SIGMA_OK;
}

extern "C" uint32_t sigma_num_rounds(const sigma_ctx* c) { return c ? c->nRounds : 0; }
extern "C" int sigma_round_reports(const sigma_ctx* c, sigma_round_report* out, uint32_t max_rounds) {
    if (!c || !out) return SIGMA_BAD_ARGUMENT;
    const u32 n = c->nRounds < max_rounds ? c->nRounds : max_rounds;
    memcpy(out, c->rounds, n * sizeof(sigma_round_report));
    return SIGMA_OK;
}

// ------------------------------------------------------------------ proof stream (host side)
extern "C" int sigma_set_proof_sink(sigma_ctx* c, sigma_proof_sink sink, void* user) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    c->proofSink = sink; c->proofUser = user;
    return SIGMA_OK;
}
extern "C" int sigma_proof_chunks(const sigma_ctx* c, uint32_t* num_chunks, uint64_t* total_bytes, uint32_t* capacity) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    if (num_chunks) *num_chunks = c->nProofChunks;
    if (total_bytes) *total_bytes = c->proofAllSize;
    if (capacity) *capacity = c->hdc ? c->hdc->proofCap : 0;
    return SIGMA_OK;
}
extern "C" int sigma_proof_chunk_size(const sigma_ctx* c, uint32_t chunk, uint64_t* num_bytes) {
    if (!c || !num_bytes || chunk >= c->nProofChunks) return SIGMA_BAD_ARGUMENT;
    *num_bytes = c->proofEnds[chunk] - (chunk ? c->proofEnds[chunk - 1] : 0);
    return SIGMA_OK;
}
extern "C" int sigma_proof_chunk_copy(const sigma_ctx* c, uint32_t chunk, uint8_t* out) {
    if (!c || !out || chunk >= c->nProofChunks) return SIGMA_BAD_ARGUMENT;
    const u64 lo = chunk ? c->proofEnds[chunk - 1] : 0;
    memcpy(out, c->proofAll + lo, c->proofEnds[chunk] - lo);
    return SIGMA_OK;
}

// ------------------------------------------------------------------ store
extern "C" int sigma_snapshot(sigma_ctx* c, uint64_t* num_clauses, uint64_t* num_literals) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    launchCount(c);
    int rc = syncCounters(c);
    if (rc) return rc;
    if (num_clauses) *num_clauses = c->hdc->liveCls;
    if (num_literals) *num_literals = c->hdc->liveLits;
    return SIGMA_OK;
}

extern "C" int sigma_result_sizes(sigma_ctx* c, uint64_t* num_clauses, uint64_t* num_literals, uint64_t* num_resolved,
                                  uint64_t* num_trail) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    // sizes only: the live counts of the last k_count are still valid after sigma_finish (nothing touched the store);
    // between rounds one counting pass is enough - the store kernels run once, in sigma_store* itself
    int rc;
    if (!(c->loopDone && c->countsFresh)) {
        launchCount(c);
        if ((rc = syncCounters(c))) return rc;
        c->countsFresh = c->loopDone;
    }
    const bool live = c->cnfstate == SIGMA_UNSOLVED;
    if (num_clauses) *num_clauses = live ? c->hdc->liveCls : 0;
    if (num_literals) *num_literals = live ? c->hdc->liveLits : 0;
    if (num_resolved) *num_resolved = c->hdc->resolvedSize;
    if (num_trail) *num_trail = c->hdc->trailSize;
    return SIGMA_OK;
}

extern "C" int sigma_store(sigma_ctx* c, uint32_t* bits, uint32_t* sig, uint64_t* offs, uint32_t* lits, uint8_t* eliminated,
                           uint32_t* resolved, uint32_t* trail) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = syncCounters(c);
    if (rc) return rc;
    u64 nc = 0, nl = 0;
    // -aggresivesort orders the final write-back only (cacheCNF); a store between two sigma_round calls is a snapshot in ref order
    if ((rc = launchStore(c, &nc, &nl, 0, c->loopDone))) return rc;
    const int dst = 1 - c->cur;
    const bool live = c->cnfstate == SIGMA_UNSOLVED;
    if (offs) {
        if (live) CUDA_TRY(cudaMemcpyAsync(offs, c->flag64, (nc + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        else offs[0] = 0;
    }
    if (live && nc) {
        const u32* oBits = (const u32*)c->hdr[dst];
        if (bits) CUDA_TRY(cudaMemcpyAsync(bits, oBits, nc * 4, cudaMemcpyDeviceToHost, c->stream));
        if (sig) CUDA_TRY(cudaMemcpyAsync(sig, oBits + c->capC, nc * 4, cudaMemcpyDeviceToHost, c->stream));
        if (lits && nl) CUDA_TRY(cudaMemcpyAsync(lits, c->pool[dst], nl * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (eliminated) CUDA_TRY(cudaMemcpyAsync(eliminated, c->eliminated, (size_t)c->V + 1, cudaMemcpyDeviceToHost, c->stream));
    if (resolved && c->hdc->resolvedSize)
        CUDA_TRY(cudaMemcpyAsync(resolved, c->resolved, (size_t)c->hdc->resolvedSize * 4, cudaMemcpyDeviceToHost, c->stream));
    if (trail && c->hdc->trailSize)
        CUDA_TRY(cudaMemcpyAsync(trail, c->trail, (size_t)c->hdc->trailSize * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SIGMA_OK;
}

// What Solver::writeBackCNF -> newClause(SCLAUSE&) actually reads (sclause.cpp:22-55): word 0, the size and the literals.
// No signatures, no 64-bit offsets: 8 + 4 |c| bytes per clause over PCIe instead of 16 + 4 |c|.
extern "C" int sigma_store_compact(sigma_ctx* c, uint32_t* bits, uint32_t* sizes, uint32_t* lits, uint8_t* eliminated,
                                   uint32_t* resolved, uint32_t* trail) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = syncCounters(c);
    if (rc) return rc;
    u64 nc = 0, nl = 0;
    if ((rc = launchStore(c, &nc, &nl, 2, c->loopDone))) return rc;
    const int dst = 1 - c->cur;
    if (c->cnfstate == SIGMA_UNSOLVED && nc) {
        const u32* oBits = (const u32*)c->hdr[dst];
        if (bits) CUDA_TRY(cudaMemcpyAsync(bits, oBits, nc * 4, cudaMemcpyDeviceToHost, c->stream));
        if (sizes) CUDA_TRY(cudaMemcpyAsync(sizes, oBits + c->capC, nc * 4, cudaMemcpyDeviceToHost, c->stream));
        if (lits && nl) CUDA_TRY(cudaMemcpyAsync(lits, c->pool[dst], nl * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (eliminated) CUDA_TRY(cudaMemcpyAsync(eliminated, c->eliminated, (size_t)c->V + 1, cudaMemcpyDeviceToHost, c->stream));
    if (resolved && c->hdc->resolvedSize)
        CUDA_TRY(cudaMemcpyAsync(resolved, c->resolved, (size_t)c->hdc->resolvedSize * 4, cudaMemcpyDeviceToHost, c->stream));
    if (trail && c->hdc->trailSize)
        CUDA_TRY(cudaMemcpyAsync(trail, c->trail, (size_t)c->hdc->trailSize * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SIGMA_OK;
}

extern "C" int sigma_store_sclauses(sigma_ctx* c, uint32_t* data_words, uint64_t* refs) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = syncCounters(c);
    if (rc) return rc;
    u64 nc = 0, nl = 0;
    if ((rc = launchStore(c, &nc, &nl, 1, c->loopDone))) return rc;
    if (c->cnfstate != SIGMA_UNSOLVED || !nc) return SIGMA_OK;
    const int dst = 1 - c->cur;
    if (data_words) CUDA_TRY(cudaMemcpyAsync(data_words, c->pool[dst], (nc * NBUCKETS + nl) * 4, cudaMemcpyDeviceToHost, c->stream));
    if (refs) CUDA_TRY(cudaMemcpyAsync(refs, c->flag64, nc * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SIGMA_OK;
}

// ------------------------------------------------------------------ trail of the device prop() calls
extern "C" int sigma_trail_info(const sigma_ctx* c, uint64_t* total, uint32_t* last_from, uint32_t* last_count, uint32_t* last_seeds) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    if (total) *total = c->hdc->trailSize;
    if (last_from) *last_from = c->lastPropTrail0;
    if (last_count) *last_count = c->lastPropTotal;
    if (last_seeds) *last_seeds = c->lastPropSeeds < c->lastPropTotal ? c->lastPropSeeds : c->lastPropTotal;
    return SIGMA_OK;
}
extern "C" int sigma_copy_trail(sigma_ctx* c, uint64_t from, uint64_t count, uint32_t* out) {
    if (!c || !c->begun) return SIGMA_NOT_LOADED;
    if (!count) return SIGMA_OK;
    if (!out || from + count > c->hdc->trailSize) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(out, c->trail + from, count * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SIGMA_OK;
}

// ------------------------------------------------------------------ pinned host buffers for the callers' edges
extern "C" void* sigma_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void sigma_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ device-resident result: views and continuation
extern "C" int sigma_device_view(sigma_ctx* c, sigma_device_cnf* v) {
    if (!c || !v) return SIGMA_BAD_ARGUMENT;
    if (!c->begun) return SIGMA_NOT_LOADED;
    memset(v, 0, sizeof *v);
    v->device = c->device; v->stream = (void*)c->stream;
    v->max_var = c->V; v->clause_slots = c->hdc->numCls; v->pool_words = c->hdc->poolUsed;
    v->live_clauses = c->numClauses; v->live_literals = c->numLiterals;
    v->headers = c->hdr[c->cur]; v->literals = c->pool[c->cur];
    v->ot_start = c->otStart; v->ot_size = c->otSize; v->ot_entries = c->occurs;
    v->eliminated = c->eliminated; v->vstate = c->vstate; v->vorg = c->vorg;
    v->elected = c->elected; v->num_elected = c->numElected;
    v->units = c->units; v->resolved = c->resolved; v->resolved_words = c->hdc->resolvedSize;
    v->trail = c->trail; v->trail_units = c->hdc->trailSize;
    return SIGMA_OK;
}

// live clauses of the finished call -> the input arrays of the next one (order kept), originals counted
__global__ void k_rebase(const uint4* __restrict__ hdr, const u32* __restrict__ pool, u32 n, const u32* __restrict__ pCls, const u32* __restrict__ pLits,
                         u32* __restrict__ inLits, u64* __restrict__ inOffs, u32* __restrict__ inMeta, const u32* totCls, const u32* totLits,
                         unsigned long long* orgCL) {
    u32 oc = 0, ol = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = hdr[i];
        if (C_DELETED(h.w)) continue;
        const u32 j = pCls[i], off = pLits[i];
        const u32* s = pool + h.x;
        for (u32 k = 0; k < h.y; k++) inLits[off + k] = s[k];
        inOffs[j] = off;
        inMeta[j] = C_LEARNT(h.w) ? (h.w & ~(CB_DELETED | CB_MOLTEN | CB_ADDED)) : 0u;
        if (!C_LEARNT(h.w)) { oc++; ol += h.y; }
    }
    oc = warpSum(oc); ol = warpSum(ol);
    if ((threadIdx.x & 31u) == 0 && oc) { atomicAdd(&orgCL[0], (unsigned long long)oc); atomicAdd(&orgCL[1], (unsigned long long)ol); }
    if (blockIdx.x == 0 && threadIdx.x == 0) inOffs[*totCls] = *totLits;
}
// variables eliminated (not forced) or assigned by the finished call are inactive in the next one; counts the active ones
__global__ void k_vstate_next(const unsigned char* __restrict__ vstate, const unsigned char* __restrict__ eliminated, u32 V,
                              unsigned char* __restrict__ vstate0, u32* active) {
    u32 a = 0;
    for (u32 v = 1 + blockIdx.x * blockDim.x + threadIdx.x; v <= V; v += gridDim.x * blockDim.x) {
        unsigned char s = vstate[v];
        const unsigned char e = eliminated[v];
        if (!s && e && !(e & FORCED_MASK)) s = 3;   // MELTED_M (markEliminated, transfer.cu:42-60)
        vstate0[v] = s;
        a += !s;
    }
    a = warpSum(a);
    if ((threadIdx.x & 31u) == 0 && a) atomicAdd(active, a);
}

extern "C" int sigma_continue(sigma_ctx* c, uint64_t num_new, const uint32_t* new_lits, const uint64_t* new_offs, const uint32_t* new_meta,
                              const uint8_t* vstate, const uint8_t* assumed) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    if (!c->begun || !c->loopDone) return SIGMA_NOT_LOADED;
    if (c->cnfstate != SIGMA_UNSOLVED) { snprintf(c->err, sizeof c->err, "sigma_continue: the formula is already decided"); return SIGMA_BAD_ARGUMENT; }
    if (num_new && (!new_lits || !new_offs)) return SIGMA_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = syncCounters(c);
    if (rc) return rc;
    const u32 n = c->hdc->numCls;
    const u64 newL = num_new ? new_offs[num_new] - new_offs[0] : 0;
    u32* tot = c->dc->scratch;
    unsigned long long* orgCL = (unsigned long long*)c->dc->orgCL;
    CUDA_TRY(cudaMemsetAsync(orgCL, 0, 16, c->stream));
    CUDA_TRY(cudaMemsetAsync(tot, 0, 8, c->stream));
    u64 nc = 0, nl = 0;
    if (n) {
        // same selection as the store: flags, two scans; then the copy goes into the input arrays instead of the staging buffers
        if ((rc = launchStore(c, &nc, &nl, 3, false))) return rc;
        if (nc + num_new > c->inCapC || nl + newL > c->inCapL) {
            snprintf(c->err, sizeof c->err, "sigma_continue: %llu clauses / %llu literals do not fit the input arrays carved at sigma_load (%llu / %llu): load again",
                     (unsigned long long)(nc + num_new), (unsigned long long)(nl + newL), (unsigned long long)c->inCapC, (unsigned long long)c->inCapL);
            return SIGMA_CNFALLOC_FAIL;
        }
        LAUNCH(c, k_rebase, gridFor(n, 256), 256, 0, c->hdr[c->cur], c->pool[c->cur], n, c->flagA, c->flagB, c->inLits, c->inOffs, c->inMetaBuf, tot, tot + 1, orgCL);
    } else CUDA_TRY(cudaMemsetAsync(c->inOffs, 0, 8, c->stream));
    unsigned long long horg[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(horg, orgCL, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    u64 orgC = horg[0], orgL = horg[1];
    // the clauses the host learnt since (or any other delta), appended in the order given
    if (num_new) {
        u64* shifted = (u64*)malloc((num_new + 1) * sizeof(u64));
        if (!shifted) return SIGMA_AWAKEN_FAIL;
        for (u64 i = 0; i <= num_new; i++) shifted[i] = nl + (new_offs[i] - new_offs[0]);
        for (u64 i = 0; i < num_new; i++) if (!new_meta || !(new_meta[i] & CB_LEARNT)) { orgC++; orgL += new_offs[i + 1] - new_offs[i]; }
        cudaError_t e = cudaMemcpyAsync(c->inOffs + nc, shifted, (num_new + 1) * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->inLits + nl, new_lits + new_offs[0], newL * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = new_meta ? cudaMemcpyAsync(c->inMetaBuf + nc, new_meta, num_new * 4, cudaMemcpyHostToDevice, c->stream)
                                           : cudaMemsetAsync(c->inMetaBuf + nc, 0, num_new * 4, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        free(shifted);
        if (e != cudaSuccess) { snprintf(c->err, sizeof c->err, "sigma_continue: %s", cudaGetErrorString(e)); return -(int)e; }
    }
    // per-variable state of the next call
    const size_t V1 = (size_t)c->V + 1;
    if (vstate) {
        CUDA_TRY(cudaMemcpyAsync(c->vstate0, vstate, V1, cudaMemcpyHostToDevice, c->stream));
        i64 un = c->V;
        for (size_t v = 1; v < V1; v++) if (vstate[v]) un--;
        c->unassigned0 = un;
    } else {
        u32* active = &c->dc->scratch[14];
        CUDA_TRY(cudaMemsetAsync(active, 0, 4, c->stream));
        LAUNCH(c, k_vstate_next, gridFor(c->V, 256), 256, 0, c->vstate, c->eliminated, c->V, c->vstate0, active);
        u32 ha = 0;
        CUDA_TRY(cudaMemcpyAsync(&ha, active, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->unassigned0 = ha;
    }
    if (assumed) {
        if (!c->assumedBuf) { snprintf(c->err, sizeof c->err, "sigma_continue: no assumption buffer"); return SIGMA_BAD_ARGUMENT; }
        c->assumed = c->assumedBuf;
        CUDA_TRY(cudaMemcpyAsync(c->assumed, assumed, V1, cudaMemcpyHostToDevice, c->stream));
    } else c->assumed = nullptr;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // the formula of the next call; a later inprocessing call (stats.sigma.calls > 1: counts originals only, bounded.cuh:428-430)
    c->C0 = nc + num_new; c->L0 = nl + newL;
    c->orgClauses = orgC; c->orgLiterals = orgL;
    c->inMeta = c->inMetaBuf;
    c->o.sigma_calls++;
    u64 lc = 0, lw = 0;
    if (!logicalCaps(c, c->C0, c->L0, orgC, orgL, &lc, &lw) || lc > c->capC || lw > c->capW) {
        snprintf(c->err, sizeof c->err, "sigma_continue: the logical capacities of the continued formula exceed the arena: load again");
        c->needReload = true;
        return SIGMA_CNFALLOC_FAIL;
    }
    c->logC = lc; c->logW = lw;
    const u64 rcap = c->C0 + c->L0;
    if (rcap > c->resolvedCapPhys) { snprintf(c->err, sizeof c->err, "sigma_continue: witness stack capacity"); c->needReload = true; return SIGMA_CNFALLOC_FAIL; }
    c->resolvedCap = (u32)rcap;
    c->begun = false;
    return SIGMA_OK;
}

// ------------------------------------------------------------------ debugging / stats
extern "C" int sigma_debug_elected(sigma_ctx* c, uint32_t* out, uint32_t* n) {
    if (!c || !c->begun || !n) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    *n = c->lastElectedCount;
    // note: after BVE the array holds the survivors in its first numElected entries
    if (out && c->numElected) CUDA_TRY(cudaMemcpy(out, c->elected, (size_t)c->numElected * 4, cudaMemcpyDeviceToHost));
    *n = c->numElected;
    return SIGMA_OK;
}
extern "C" int sigma_debug_hist(sigma_ctx* c, uint32_t* out) {
    if (!c || !c->begun || !out) return SIGMA_NOT_LOADED;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpy(out, c->hist, (size_t)c->ND * 4, cudaMemcpyDeviceToHost));
    return SIGMA_OK;
}
extern "C" int sigma_memory(const sigma_ctx* c, uint64_t* arena_bytes, uint64_t* peak_used, uint64_t* cuda_mallocs) {
    if (!c) return SIGMA_BAD_ARGUMENT;
    if (arena_bytes) *arena_bytes = c->arenaBytes;
    if (peak_used) *peak_used = c->arenaPeak;
    if (cuda_mallocs) *cuda_mallocs = c->cudaMallocs;
    return SIGMA_OK;
}

// ------------------------------------------------------------------ stage entry points
extern "C" int sigma_stage_prep(int device, uint64_t num_clauses, uint32_t* lits, const uint64_t* offs, uint32_t* sig) {
    if (!lits || !offs) return SIGMA_BAD_ARGUMENT;
    sigma_ctx* c = nullptr;
    int rc = sigma_create(device, nullptr, &c);
    if (rc) return rc;
    u32 maxv = 1;
    const u64 L = offs[num_clauses];
    for (u64 i = 0; i < L; i++) if ((lits[i] >> 1) > maxv) maxv = lits[i] >> 1;
    rc = sigma_load(c, maxv, num_clauses, lits, offs, nullptr, nullptr, nullptr, nullptr);
    if (!rc) rc = sigma_begin(c);
    if (!rc && num_clauses) {
        // headers carry the signatures; literals stay at their input offsets
        uint4* h = (uint4*)malloc(num_clauses * sizeof(uint4));
        cudaError_t e = cudaMemcpyAsync(h, c->hdr[c->cur], num_clauses * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(lits, c->pool[c->cur], L * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = -(int)e;
        else if (sig) for (u64 i = 0; i < num_clauses; i++) sig[i] = h[i].z;
        free(h);
    }
    sigma_destroy(c);
    return rc;
}

__global__ void k_stage_hist(const u32* __restrict__ lits, u64 n, u32* __restrict__ hist, u32 nbins, u32* bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 l = lits[i];
        if (l < nbins) atomicAdd(&hist[l], 1u); else *bad = 1u;
    }
}
extern "C" int sigma_stage_histogram(int device, uint64_t num_lits, const uint32_t* lits, uint32_t nbins, uint32_t* hist) {
    if (!lits || !hist) return SIGMA_BAD_ARGUMENT;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return -(int)e;
    u32 *dl = nullptr, *dh = nullptr;
    if ((e = cudaMalloc(&dl, num_lits * 4 + 4)) != cudaSuccess) return -(int)e;
    if ((e = cudaMalloc(&dh, (size_t)nbins * 4 + 8)) != cudaSuccess) { cudaFree(dl); return -(int)e; }
    cudaMemcpy(dl, lits, num_lits * 4, cudaMemcpyHostToDevice);
    cudaMemset(dh, 0, (size_t)nbins * 4 + 8);
    if (num_lits) k_stage_hist<<<gridFor(num_lits, 256), 256>>>(dl, num_lits, dh, nbins, dh + nbins);
    e = cudaMemcpy(hist, dh, (size_t)nbins * 4, cudaMemcpyDeviceToHost);
    u32 bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, dh + nbins, 4, cudaMemcpyDeviceToHost);
    cudaFree(dl); cudaFree(dh);
    if (e != cudaSuccess) return -(int)e;
    return bad ? SIGMA_BAD_ARGUMENT : SIGMA_OK;   // a literal >= nbins
}
