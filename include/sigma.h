/* include/sigma.h -- C ABI of the B200-native SIGmA inprocessing engine (libsigma_b200.so).
 *
 * Drop-in boundary for ParaFROST's GPU simplifier.  The reference has no plugin API: the
 * simplifier is a set of `Solver` members compiled into libparafrost.a
 * (src/gpu/solver.hpp:674-789).  A shim translation unit re-defines the 7 symbols the host
 * objects import from the reference's CUDA objects (INTEGRATION.md shows it)
 * and forwards to the entry points below, each of which replaces the reference interface
 * named beside it.  Plain pointers and sizes only; no exceptions, no exit(): every call
 * returns 0 on success, a positive reference `simpstate` code
 * (src/gpu/constants.cuh:28-31) or a negative CUDA error (-cudaError_t).
 *
 * Re-entrant and context based (the reference is process-global): one context = one CNF on
 * one device; contexts on different devices/threads are independent (instance-parallel
 * batches, no collective).
 *
 * Literal encoding is the reference's: lit = 2*var + sign, var >= 1 (constants.hpp:72-80).
 */
#ifndef SIGMA_B200_H
#define SIGMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGMA_OK             0
#define SIGMA_AWAKEN_FAIL    1   /* AWAKEN_FAIL    constants.cuh:29 */
#define SIGMA_CNFALLOC_FAIL  2   /* CNFALLOC_FAIL  constants.cuh:30 */
#define SIGMA_OTALLOC_FAIL   3   /* OTALLOC_FAIL   constants.cuh:31 */
#define SIGMA_BAD_ARGUMENT   16
#define SIGMA_NOT_LOADED     17
#define SIGMA_OVERFLOW       18  /* a device vector (units / resolved) would overflow; the reference only asserts (vector.cu:77-98) */

#define SIGMA_UNSAT    0         /* CNFState, constants.hpp:27 */
#define SIGMA_SAT      1
#define SIGMA_UNSOLVED 2

/* Simplifier options: the reference's flags with their defaults
 * (src/gpu/options.cpp:24-43, src/gpu/options.cu:36-60); replaces OPTION/GOPTION/kOpts. */
typedef struct sigma_opts {
    int32_t  phases;            /* --phases=5 */
    int32_t  ve_en;             /* -ve */
    int32_t  ve_plus_en;        /* -veextend */
    int32_t  sub_en;            /* -sub */
    int32_t  bce_en;            /* -bce (off) */
    int32_t  ere_en;            /* -ere */
    int32_t  all_en;            /* -all */
    uint32_t mu_pos, mu_neg;    /* --mupos/--muneg = 32 */
    uint32_t lcve_min_vars;     /* --electionsmin = 2 */
    uint32_t lcve_max_occurs;   /* --electionsmax = 3000 */
    int32_t  lcve_clause_max;   /* --lcveclausemax = 30000 */
    int32_t  phase_lits_min;    /* --eliminatedlitsmin = 500 */
    int32_t  shrink_rate;       /* --collectfreq = 2 */
    double   lits_mul;          /* --literalsmul = 1.0 */
    int32_t  ve_fun_en;         /* -vefunction */
    int32_t  ve_lbound_en;      /* -velitsbound (off) */
    uint32_t ve_clause_max;     /* --resolventmax = 100 */
    uint32_t xor_max_arity;     /* --xormaxarity = 10 */
    int32_t  ere_clause_max;    /* min(--ereclausemax = 250, 250) */
    uint32_t ere_max_occurs;    /* --eremaxoccurs = 3000 */
    uint32_t sub_max_occurs;    /* --submaxoccurs = 3000 */
    uint32_t bce_max_occurs;    /* --bcemaxoccurs = 3000 */
    uint32_t sh_max_bve_out1;   /* SH_MAX_BVE_OUT1 of the replaced build: 250 (EXTSHMEM) / 190 */
    int32_t  sigma_calls;       /* stats.sigma.calls of this call: 1 = preprocessing */
    int32_t  final_gc;          /* reserved, not read: every store entry point and sigma_continue compact on the way out, so the
                                   reallocCNF(true) of simplify(skip_transfer_to_host) (simplify.cu:224-226) has no counterpart to switch */
    int32_t  profile;           /* -profilegpu: per-stage CUDA-event times */
    int32_t  aggr_cnf_sort;     /* -aggresivesort (off): clauses leave in OLIST_CMP order (cnf.cu:232-233, key.cuh:67-83) */
    int32_t  proof_en;          /* -proof (off): device DRAT stream (proof.cu, proofutils.cuh); set before sigma_load */
    int32_t  lcve_fast;         /* -lcvefast: the reference CLI's default election (options.cpp:32, lcve.cu:150-217,338-366) - a maximal
                                   independent set over the FILTERED candidates (the stop conditions of the serial walk become
                                   filters, so more variables are elected).  sigma_default_opts leaves it 0 = -no-lcvefast, the
                                   deterministic mode every parity run uses; the shim forwards the CLI's value.  The elected SET
                                   equals the reference's; its order there comes from atomics, here it is the rank order. */
    int32_t  log_reductions;    /* --verbose>=2 of the reference: LOGREDALL / LOGREDCL (logging.hpp:152-158, count.cu:185-208) - live
                                   counts after BCP, SUB, BVE, BCE and ERE of every round (one counting pass + read-back per stage),
                                   fetched with sigma_reduction_log */
} sigma_opts;

/* Per-round report; replaces the LOG2 lines + inf.* updates of simplify.cu:163-186. */
typedef struct sigma_round_report {
    uint32_t round;             /* phase index of this iteration */
    uint32_t kind;              /* 0 = SUB/BVE/BCE round, 1 = ERE-and-stop, 2 = loop left before eliminations */
    uint32_t elected;           /* vars->numElected after LCVE */
    uint32_t eliminated;        /* inf.currDeletedVars */
    uint32_t resolvents;        /* clauses appended by BVE */
    uint32_t units;             /* vars->nUnits produced this round */
    uint32_t propagated;        /* units propagated by prop() at the top of this round */
    uint32_t gc;                /* 1 if the CNF was compacted this round */
    uint64_t clauses;           /* inf.numClauses after the round */
    uint64_t literals;          /* inf.numLiterals after the round */
    uint64_t literals_in;       /* live literals when the round started */
    float    ms;                /* wall ms of the round (host clock, stream synchronised) */
    uint32_t trail_added;       /* trail entries appended by this round's prop(): the first `propagated` are the BVE/SUB units */
} sigma_round_report;

/* Whole-call report; replaces stats.sigma.* (statistics.hpp:33-35). */
typedef struct sigma_report {
    int32_t  cnfstate;          /* SIGMA_UNSAT / SAT / UNSOLVED */
    int32_t  simpstate;
    uint32_t rounds;
    uint32_t eliminated_vars;   /* vars->currMelted */
    uint64_t clauses, literals; /* live after the call */
    uint64_t clauses_in, literals_in;
    uint64_t resolved_words;
    uint64_t trail_units;
    double   ms_total;          /* awaken .. end of loop, host clock */
    /* -profilegpu stage totals in ms (statistics.cpp:37-49): vo sig io gc cot sot rot ve sub bce ere + prop lcve */
    float    stage_ms[16];
    uint64_t kernel_launches;   /* kernels launched by this call */
    double   ms_device;         /* sigma_run only: CUDA-event time begin..end on the launch stream */
} sigma_report;

typedef struct sigma_ctx sigma_ctx;

/* options.cpp:168-300 / options.cu:66-89 */
void sigma_default_opts(sigma_opts* o);
void sigma_normalize_opts(sigma_opts* o);          /* derivations of options.cpp:291-296 */

/* Solver::optSimp + createStreams (simplify.cu:243, solver.hpp:728): bind a device, create the
 * stream; no device memory yet. */
int  sigma_create(int device, const sigma_opts* o, sigma_ctx** out);
/* Solver::freeSimp (simplify.cu:254) */
int  sigma_destroy(sigma_ctx* c);
int  sigma_set_opts(sigma_ctx* c, const sigma_opts* o);
/* streams[] of the reference (solver.hpp:728): run every launch and copy of this context on the
 * caller's CUDA stream (a cudaStream_t passed as void*), so the caller can order and time it. */
int  sigma_set_stream(sigma_ctx* c, void* cuda_stream);

/* Solver::awaken's host half: extractCNF + reflectCNF (cnf.cu:166-184) and cuMM::init*/
/* (memory.cu:99-387).  HOST buffers in CSR form; sizes the arena (ONE cudaMalloc, reused
 * while it fits) and copies the formula to the device.
 *   lits[offs[C]]  literals, offs[C+1]
 *   meta[C]        NULL or per clause: bit0 learnt, bits 4..5 usage, bits 6.. lbd (SCLAUSE word 0)
 *   vorg[V+1]      NULL (identity) or current->original variable map   (solver.hpp:84)
 *   vstate[V+1]    NULL or sp->vstate[].state: non-zero = inactive      (vstate.hpp:26-29)
 *   assumed[V+1]   NULL or the incremental assumption mask               (lcve.cu:88,163) */
int  sigma_load(sigma_ctx* c, uint32_t max_var, uint64_t num_clauses,
                const uint32_t* lits, const uint64_t* offs, const uint32_t* meta,
                const uint32_t* vorg, const uint8_t* vstate, const uint8_t* assumed);

/* sigma_load with 32-bit clause offsets (offs32[num_clauses] < 2^32 literals): 4 instead of 8 bytes per clause over PCIe;
 * the offsets are widened on the device.  Same result as sigma_load. */
int  sigma_load32(sigma_ctx* c, uint32_t max_var, uint64_t num_clauses,
                  const uint32_t* lits, const uint32_t* offs32, const uint32_t* meta,
                  const uint32_t* vorg, const uint8_t* vstate, const uint8_t* assumed);

/* The same from the reference's own host mirror `hcnf` (CNF::newClause, cnf.cuh:82-97): the SCLAUSE
 * record stream {word 0 = st:2 f:1 a:1 u:2 lbd:26, sig, size, literals...} of num_words words and the
 * uint64 word offsets `refs[num_clauses]`, gap-free in ref order - exactly the two buffers
 * Solver::reflectCNF copies to the device (cnf.cu:166-174); unpacked on the device.  The inverse of
 * sigma_store_sclauses. */
int  sigma_load_sclauses(sigma_ctx* c, uint32_t max_var, uint64_t num_clauses, const uint32_t* data_words,
                         uint64_t num_words, const uint64_t* refs, const uint32_t* vorg, const uint8_t* vstate,
                         const uint8_t* assumed);

/* Solver::simplifying (simplify.cu:136-241): awaken's device half (prep_cnf_k) + the round
 * loop.  Can be called repeatedly on the loaded formula (each call restarts from it). */
int  sigma_run(sigma_ctx* c, sigma_report* rep);
/* The same, one loop iteration at a time: sigma_begin, then sigma_round until *done. */
int  sigma_begin(sigma_ctx* c);
int  sigma_round(sigma_ctx* c, sigma_round_report* rep, int* done);
int  sigma_finish(sigma_ctx* c, sigma_report* rep);
uint32_t sigma_num_rounds(const sigma_ctx* c);
int  sigma_round_reports(const sigma_ctx* c, sigma_round_report* out, uint32_t max_rounds);

/* Solver::cacheCNF / cacheResolved / cacheEliminated / cacheUnits (cnf.cu:200-237,
 * transfer.cu:62-97): sizes first, then the device->host copy into HOST buffers.
 * The clause stream is the reference's: clauses in ref order, bits = SCLAUSE word 0. */
int  sigma_result_sizes(sigma_ctx* c, uint64_t* num_clauses, uint64_t* num_literals,
                        uint64_t* num_resolved, uint64_t* num_trail);
int  sigma_store(sigma_ctx* c, uint32_t* bits, uint32_t* sig, uint64_t* offs, uint32_t* lits,
                 uint8_t* eliminated, uint32_t* resolved, uint32_t* trail);
/* What writeBackCNF -> newClause(SCLAUSE&) reads (cnf.cu:186-198, sclause.cpp:22-55): word 0, size and literals of every
 * clause in ref order - no signatures, no 64-bit offsets (8 + 4|c| bytes per clause over PCIe instead of 16 + 4|c|). */
int  sigma_store_compact(sigma_ctx* c, uint32_t* bits, uint32_t* sizes, uint32_t* lits,
                         uint8_t* eliminated, uint32_t* resolved, uint32_t* trail);
/* the reference's own record stream {bits, sig, size, lits...} + uint64 refs, for newClause(SCLAUSE&) */
int  sigma_store_sclauses(sigma_ctx* c, uint32_t* data_words, uint64_t* refs);

/* Units assigned by the device prop() calls (elimbcp.cu:144-215), in the order the reference's host loop enqueues them
 * (elimbcp.cu:185-200): per prop() first the units SUB/BVE produced (enqueueDevUnit), then the derived ones (enqueueUnit).
 * sigma_trail_info: size of the whole trail and the range / seed count of the LAST prop(); valid between sigma_round calls
 * and inside the proof sink, so a host can enqueue a round's units before that round's proof chunk, as the reference does. */
int  sigma_trail_info(const sigma_ctx* c, uint64_t* total, uint32_t* last_from, uint32_t* last_count, uint32_t* last_seeds);
int  sigma_copy_trail(sigma_ctx* c, uint64_t from, uint64_t count, uint32_t* out);

/* Pinned (page-locked) host memory for the callers' edges: extractCNF / writeBackCNF buffers (cnf.cu:176-198) reach PCIe
 * speed only from pinned pages; replaces cuMM::createMirror's pinned mirror (memory.cu). */
void* sigma_pinned_alloc(size_t bytes);
void  sigma_pinned_free(void* p);

/* Device-resident result (simplify(skip_transfer_to_host), simplify.cu:221-229): views of what stays in the context's
 * arena after sigma_run - replaces Solver::getDeviceCNF / getDeviceOT / getVars (solver.hpp:694-705) for integrations
 * that keep working on the GPU (gpu4bmc).  Pointers are DEVICE pointers, valid until the next sigma_load / sigma_begin /
 * sigma_destroy; clause i = headers[i] {x: offset into literals, y: size, z: signature, w: SCLAUSE word 0 (bit 1 =
 * deleted)}, i < clause_slots. */
typedef struct sigma_device_cnf {
    int32_t  device;
    void*    stream;            /* cudaStream_t all work of the context is ordered on */
    uint32_t max_var;
    uint32_t clause_slots;      /* cnf->size(): live and deleted slots */
    uint64_t pool_words;
    uint64_t live_clauses, live_literals;
    const void*     headers;    /* uint4[clause_slots] */
    const uint32_t* literals;
    const uint32_t* ot_start;   /* [2V+3] occurrence lists: entries ot_entries[ot_start[lit] .. + ot_size[lit]) */
    const uint32_t* ot_size;
    const uint32_t* ot_entries;
    const uint8_t*  eliminated; /* [V+1] MELTING 1 | ADDING 2 | FORCED 4 (constants.cuh:33-36) */
    const uint8_t*  vstate;     /* [V+1] */
    const uint32_t* vorg;       /* [V+1] */
    const uint32_t* elected;    /* survivors of the last election */
    uint32_t num_elected;
    const uint32_t* units;
    const uint32_t* resolved;   /* witness stack of this call */
    uint64_t resolved_words;
    const uint32_t* trail;
    uint64_t trail_units;
} sigma_device_cnf;
int  sigma_device_view(sigma_ctx* c, sigma_device_cnf* out);

/* The next inprocessing call ON the resident result: the live clauses of the finished call become the input of the next
 * one without crossing PCIe (device-side copy into the input arrays, order kept, learnt flags / lbd / usage kept);
 * only the clauses the host added since (num_new: learnt clauses, ...) and - optionally - the per-variable state travel.
 *   vstate  NULL: derived on the device (inactive = inactive before, eliminated or assigned by the finished call)
 *   assumed NULL: no assumptions
 * opts.sigma_calls is advanced by one (later calls count original clauses only, bounded.cuh:428-430).  Fetch the witness
 * stack / eliminated / trail of the finished call first (sigma_store_compact with NULL clause buffers): they restart.
 * Then sigma_run / sigma_begin as usual.  SIGMA_CNFALLOC_FAIL if the continued formula outgrew what sigma_load carved. */
int  sigma_continue(sigma_ctx* c, uint64_t num_new, const uint32_t* new_lits, const uint64_t* new_offs, const uint32_t* new_meta,
                    const uint8_t* vstate, const uint8_t* assumed);

/* Per-stage reduction tables (LOGREDALL "BCP / BVE Reductions", LOGREDCL "SUB / BCE / ERE Reductions"; elimbcp.cu:203,
 * elimination.cu:245,276,289,304): with opts.log_reductions one entry per stage that ran, in execution order. */
typedef struct sigma_stage_reduction {
    uint32_t round;
    uint32_t stage;             /* 0 BCP, 1 SUB, 2 BVE, 3 BCE, 4 ERE */
    uint32_t vars_removed;      /* BCP: variables forced by prop(); BVE: variables eliminated; else 0 */
    uint32_t pad;
    uint64_t clauses_before, literals_before;   /* inf.numClauses / numLiterals when the round started */
    uint64_t clauses, literals;                 /* survived after the stage */
} sigma_stage_reduction;
int  sigma_reduction_log(const sigma_ctx* c, sigma_stage_reduction* out, uint32_t* n);   /* *n: capacity in, entries out */

/* Device DRAT proof stream; replaces cuPROOF (src/gpu/proof.cuh:30-71, proof.cu) and the proof hooks of
 * sub_k / ve_k_1 / ve_k_2 / bce_k / ere_k (proofutils.cuh).  With opts.proof_en the kernels append binary
 * DRAT lines - 'a' | 'd', the ORIGINAL literals (vorg) as 7-bit varints, 0 - to a device buffer of the
 * reference's capacity (1.5 x the proof bytes of the input literals, simplify.cu:128-132; exceeding it is
 * SIGMA_OVERFLOW where the reference only asserts).  Lines: strengthened clauses added and subsumed ones
 * deleted (SUB), substituted clauses and resolvents added (BVE), blocked (BCE) and redundant (ERE) clauses
 * deleted.  Once per round - where the reference runs cacheProof + writeProof (simplify.cu:174-184,
 * elimination.cu:305-306) - the stream is copied to pinned host memory as one chunk, handed to the sink
 * (the shim forwards it byte by byte to PROOF::write, proof.cu:185-190) and kept until the next sigma_begin.
 * The order of the lines of different variables inside a stage is unspecified, as in the reference
 * (threads reserve space with cuVecB::jump). */
typedef void (*sigma_proof_sink)(void* user, const uint8_t* bytes, uint64_t num_bytes);
int  sigma_set_proof_sink(sigma_ctx* c, sigma_proof_sink sink, void* user);
int  sigma_proof_chunks(const sigma_ctx* c, uint32_t* num_chunks, uint64_t* total_bytes, uint32_t* capacity);
int  sigma_proof_chunk_size(const sigma_ctx* c, uint32_t chunk, uint64_t* num_bytes);
int  sigma_proof_chunk_copy(const sigma_ctx* c, uint32_t chunk, uint8_t* out);

/* parity / debugging */
int  sigma_snapshot(sigma_ctx* c, uint64_t* num_clauses, uint64_t* num_literals);   /* sizes of the live CNF now */
int  sigma_debug_elected(sigma_ctx* c, uint32_t* out, uint32_t* n);                 /* elected vars of the last round */
int  sigma_debug_hist(sigma_ctx* c, uint32_t* out);                                 /* [2V+2] of the last OT build */

/* Per-kernel device times (replaces -profilegpu's cuTIMER pairs, src/gpu/timer.cuh:27-59, at kernel
 * granularity): CUDA-event pairs recorded on the context's launch stream around every kernel.
 * enable: 0 = off, 1 = on and reset the totals, 2 = on and keep the totals.
 * sigma_kernel_times: names is n*64 chars, *n = capacity in, entries out. */
int  sigma_kernel_profile(sigma_ctx* c, int enable);
int  sigma_kernel_times(sigma_ctx* c, char* names, float* ms, uint32_t* counts, uint32_t* n);
/* the same plus the ALGORITHMIC bytes each kernel name had to move over its launches (16-byte clause
 * headers, 4-byte literals and list entries; the per-variable kernels sum (4 + 16 + 4|c|) over the clauses of the
 * variables they were given, counted on the device) - the numerator of the per-kernel roofline. */
int  sigma_kernel_stats(sigma_ctx* c, char* names, float* ms, uint32_t* counts, double* bytes, uint32_t* n);

/* arena statistics (replaces cuArena's gpu_peak_used, simplify.cu:219-220) */
int  sigma_memory(const sigma_ctx* c, uint64_t* arena_bytes, uint64_t* peak_used, uint64_t* cuda_mallocs);
const char* sigma_last_error(const sigma_ctx* c);
const char* sigma_version(void);

/* Stage entry points (single kernels behind the C ABI, for parity tests and the roofline bench). */
/* prep_cnf_k (cnf.cu:45): sort literals of every clause, compute signatures; host in/out */
int  sigma_stage_prep(int device, uint64_t num_clauses, uint32_t* lits, const uint64_t* offs, uint32_t* sig);
/* flattenCNF + histSimp (cnf.cu:152, histogram.cu:54): literal histogram; host in/out */
int  sigma_stage_histogram(int device, uint64_t num_lits, const uint32_t* lits, uint32_t nbins, uint32_t* hist);

#ifdef __cplusplus
}
#endif
#endif
