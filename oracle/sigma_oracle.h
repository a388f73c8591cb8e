/* oracle/sigma_oracle.h -- TEST INFRASTRUCTURE (CPU restatement of the reference algorithm).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (parafrost_b200/, include/sigma.h) never links or calls it.
 *
 * The oracle restates, sequentially and on the host, what ParaFROST's GPU simplifier
 * (src/gpu/simplify.cu:136-241 and the kernels it launches) computes in its fixed-order mode
 * (-no-lcvefast).  It is pinned against dumps of the unmodified reference GPU binary
 * (tests/golden/, produced by oracle/ref/ref_driver.cpp on a B200).
 */
#ifndef SIGMA_ORACLE_H
#define SIGMA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors the reference's simplifier flags and their defaults
 * (src/gpu/options.cpp:24-43, src/gpu/options.cu:36-60, src/gpu/constants.cuh:63-74). */
typedef struct oracle_opts {
    int32_t  phases;            /* --phases=5 */
    int32_t  ve_en;             /* -ve (|| veextend, options.cpp:291) */
    int32_t  ve_plus_en;        /* -veextend */
    int32_t  sub_en;            /* -sub */
    int32_t  bce_en;            /* -bce (off) */
    int32_t  ere_en;            /* -ere */
    int32_t  all_en;            /* -all */
    uint32_t mu_pos, mu_neg;    /* 32, 32 */
    uint32_t lcve_min_vars;     /* electionsmin 2 */
    uint32_t lcve_max_occurs;   /* electionsmax 3000 */
    int32_t  lcve_clause_max;   /* lcveclausemax 30000 */
    int32_t  phase_lits_min;    /* eliminatedlitsmin 500 */
    int32_t  shrink_rate;       /* collectfreq 2 */
    double   lits_mul;          /* literalsmul 1.0 */
    int32_t  ve_fun_en;         /* -vefunction */
    int32_t  ve_lbound_en;      /* -velitsbound (off) */
    uint32_t ve_clause_max;     /* resolventmax 100 */
    uint32_t xor_max_arity;     /* xormaxarity 10 */
    int32_t  ere_clause_max;    /* min(ereclausemax 250, SH_MAX_ERE_OUT) */
    uint32_t ere_max_occurs;    /* 3000 */
    uint32_t sub_max_occurs;    /* 3000 */
    uint32_t bce_max_occurs;    /* 3000 */
    uint32_t sh_max_bve_out1;   /* 250 with EXTSHMEM, 190 without (gate eligibility) */
    int32_t  sigma_calls;       /* stats.sigma.calls: 1 = preprocessing call */
    int32_t  final_gc;          /* 1: compact at the end like simplify(skip_transfer_to_host) */
    int32_t  aggr_cnf_sort;     /* -aggresivesort: refs stable-sorted by OLIST_CMP before the write-back (cnf.cu:232-233) */
    int32_t  lcve_fast;         /* -lcvefast (the reference CLI's default; 0 here = the deterministic parity mode): election as a
                                   maximal independent set over the FILTERED candidates (lcve.cu:150-217, 338-366) */
} oracle_opts;

void oracle_default_opts(oracle_opts* o);
/* applies the derivations of src/gpu/options.cpp:291-296 */
void oracle_normalize_opts(oracle_opts* o);

typedef struct oracle_ctx oracle_ctx;

/* meta[i]: bit0 = learnt, bits 2..3 usage, bits 6.. lbd (same packing as SCLAUSE word 0).
 * vorg may be NULL (identity), vstate may be NULL (all active). */
int  oracle_create(const oracle_opts* o, uint32_t max_var, uint64_t num_clauses,
                   const uint32_t* lits, const uint64_t* offs, const uint32_t* meta,
                   const uint32_t* vorg, const uint8_t* vstate, oracle_ctx** out);
/* runs simplifying(); returns cnfstate (0 UNSAT, 1 SAT, 2 UNSOLVED) */
int  oracle_run(oracle_ctx* c);
/* rounds actually executed and, per round r < rounds: elected, eliminated, resolvents added,
 * clauses, literals (5 x uint64 per round) */
int  oracle_rounds(const oracle_ctx* c);
void oracle_round_stats(const oracle_ctx* c, uint64_t* out5xR);

/* result, in the reference's dump form (tests/sgd.py) */
uint64_t oracle_num_clauses(const oracle_ctx* c);
uint64_t oracle_num_literals(const oracle_ctx* c);
uint64_t oracle_num_resolved(const oracle_ctx* c);
uint64_t oracle_num_trail(const oracle_ctx* c);
void oracle_copy_result(const oracle_ctx* c, uint32_t* bits, uint32_t* sig, uint64_t* offs,
                        uint32_t* lits, uint8_t* eliminated, uint32_t* resolved, uint32_t* trail);
/* snapshot of the live clauses after round r (only kept when keep_snapshots != 0) */
void oracle_keep_snapshots(oracle_ctx* c, int keep);
/* Device DRAT stream (-proof: src/gpu/proof.cu, proofutils.cuh): enable before oracle_run.  One chunk
 * per cacheProof/writeProof pair (a SUB/BVE/BCE round, simplify.cu:174-184; the ERE round,
 * elimination.cu:305-306), in binary DRAT: 'a'|'d', 7-bit varints of the ORIGINAL literals, 0.
 * Also activates the proof guards of the BVE counting functions (resolve.cuh:66-70). */
void oracle_enable_proof(oracle_ctx* c, int on);
int  oracle_proof_chunks(const oracle_ctx* c);
uint64_t oracle_proof_chunk_size(const oracle_ctx* c, int i);
void oracle_copy_proof_chunk(const oracle_ctx* c, int i, uint8_t* out);
uint32_t oracle_proof_capacity(const oracle_ctx* c);   /* 1.5 x proof bytes of the input literals (simplify.cu:128-132) */
/* assumed[max_var+1]: variables under assumption (incremental mode) are never elected; NULL clears */
void oracle_set_assumed(oracle_ctx* c, const uint8_t* assumed);
uint64_t oracle_snapshot_clauses(const oracle_ctx* c, int round);
uint64_t oracle_snapshot_literals(const oracle_ctx* c, int round);
void oracle_copy_snapshot(const oracle_ctx* c, int round, uint32_t* bits, uint32_t* sig,
                          uint64_t* offs, uint32_t* lits);
/* elected variables of the i-th election (kept with keep_snapshots), in election order */
int  oracle_num_elections(const oracle_ctx* c);
uint64_t oracle_election_size(const oracle_ctx* c, int i);
void oracle_copy_election(const oracle_ctx* c, int i, uint32_t* out);
int  oracle_write_dump(const oracle_ctx* c, const char* path);
void oracle_destroy(oracle_ctx* c);

/* Stage-level entry points used to check single kernels of the CUDA engine. */
/* per-clause literal sort + 32-bit signature (src/gpu/cnf.cu:45-53) */
void oracle_prep(uint64_t num_clauses, uint32_t* lits, const uint64_t* offs, uint32_t* sig);
/* literal histogram over clauses (src/gpu/histogram.cu:54-72) */
void oracle_histogram(uint64_t num_lits, const uint32_t* lits, uint32_t nbins, uint32_t* hist);

/* model extension over the witness stack (src/gpu/model.cpp:101-162); value[v] in {0,1}, 1-based;
 * resolved/trail use ORIGINAL variable numbering. Returns number of flipped variables. */
uint64_t oracle_extend_model(uint8_t* value, uint32_t max_var, const uint32_t* resolved, uint64_t n);
/* returns the number of falsified clauses of a CNF under value[] */
uint64_t oracle_check_model(const uint8_t* value, uint64_t num_clauses, const uint32_t* lits, const uint64_t* offs);

/* helpers of tests/sgd.py: boundaries of the witness-stack records `[lits..., size]` (ends[k] = one past record k; returns the
 * record count, ~0 if corrupt; ends may be NULL to count only) and FNV-1a hashes of word segments */
uint64_t oracle_record_ends(const uint32_t* resolved, uint64_t n, uint64_t* ends);
void oracle_hash_segments(const uint32_t* words, const uint64_t* starts, const uint64_t* ends, uint64_t m, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
