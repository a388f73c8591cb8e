/* oracle/sigma_oracle.cpp -- TEST INFRASTRUCTURE.  CPU restatement of ParaFROST's GPU
 * simplifier ("SIGmA", /root/reference/src/gpu) in its fixed-order mode (-no-lcvefast).
 *
 * Sequential, host-only, deliberately plain.  Every function names the reference file:line
 * it restates.  The CUDA engine under parafrost_b200/ is an independent implementation
 * (different data layout, warp-cooperative kernels); this file exists so that the engine
 * can be checked bit for bit on a box without the reference, and it is itself pinned by
 * the golden dumps of the unmodified reference GPU binary (tests/golden/).
 *
 * Policy decisions where the reference is racy:
 *  - occurrence lists before sortOT are in clause-index order (B.2);
 *  - ERE runs the elected variables sequentially in elected order (B.3);
 *  - units / resolved are appended in elected order (B.4): compare as multisets / groups.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
 */
#include "sigma_oracle.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

// src/gpu/constants.hpp:72-90
inline u32 ABS(u32 l) { return l >> 1; }
inline u32 SIGN(u32 l) { return l & 1; }
inline u32 FLIP(u32 l) { return l ^ 1; }
inline u32 V2L(u32 v) { return v << 1; }
inline u32 NEG(u32 l) { return l | 1; }
inline u32 MAPHASH(u32 l) { return 1u << (l & 31); }
inline bool IS_TAUTOLOGY(u32 a, u32 b) { return (a ^ b) == 1; }
inline bool SUBSIG(u32 a, u32 b) { return !(a & ~b); }  // elimination.cuh:35

enum { ORIGINAL = 0, LEARNT = 1, DELETED = 2 };
enum { UNSAT = 0, SAT = 1, UNSOLVED = 2 };
enum { AWAKEN_SUCC = 0, AWAKEN_FAIL = 1, CNFALLOC_FAIL = 2, OTALLOC_FAIL = 3 };
// src/gpu/constants.cuh:33-62
const uint8_t MELTING_MASK = 1, ADDING_MASK = 2, FORCED_MASK = 4;
const u32 RES_MASK = 1, AOIX_MASK = 2, CORE_MASK = 3;
const u32 ADDEDCLS_MAX = 0x3FFF, ADDEDLITS_MAX = 0xFFFF, ADDEDPROOF_MAX = 0x3FFFF;
const uint8_t PROOF_ADDED = 'a', PROOF_DELETED = 'd';   // constants.hpp:45-46
const int SH_MAX_BVE_OUT2 = 120;  // irrelevant to results (same clause either path)
const int SUB_MAX_CL_SIZE = 1000;
const u32 NBUCKETS = 3;           // sizeof(SCLAUSE)/4, sclause.cuh:206-209
const u32 USAGET3 = 1;
const int LBD_TIER1 = 2;

// src/gpu/sclause.cuh:37-42
struct Clause {
    u32 st, molten, added, usage, lbd;
    u32 sig;
    int sz;
    u64 off;  // into State::pool
    u64 ref;  // logical S_REF (word offset in the reference's data arena); order == index order
    bool original() const { return st == 0; }
    bool learnt() const { return st & LEARNT; }
    bool deleted() const { return st & DELETED; }
};

typedef std::vector<u32> OL;

struct RoundStat { u64 elected, eliminated, added, clauses, literals; };

struct Snapshot {
    std::vector<u32> bits, sig, lits;
    std::vector<u64> offs;
};

} // namespace

struct oracle_ctx {
    oracle_opts o;
    u32 V = 0;
    std::vector<Clause> cls;
    std::vector<u32> pool;
    u64 data_size = 0, data_cap = 0;  // logical, in words (cnf.cuh:25-31)
    u64 refs_cap = 0;
    std::vector<OL> ot;
    std::vector<u32> hist;
    std::vector<u32> eligible, scores, elected, units, resolved, frozenList, trail, vorg;
    std::vector<uint8_t> eliminated, vstate, frozen, assumed;   // assumed: incremental assumption mask (lcve.cu:316-323), may be empty
    const u32* varcore = nullptr;
    bool varcore_dead = false;
    int phase = 0, multiplier = 0, simpstate = AWAKEN_SUCC, cnfstate = UNSOLVED;
    u64 numClauses = 0, numLiterals = 0, orgClauses = 0, orgLiterals = 0;
    u32 nUnits = 0, numElected = 0;
    i64 unassigned = 0;
    bool compacted = false;
    std::vector<RoundStat> rstats;
    bool keep_snaps = false;
    std::vector<Snapshot> snaps;
    std::vector<std::vector<u32>> electedLog;  // elected variables of every LCVE call, in order
    // BVE phase arrays (elimination.cu:146-149)
    std::vector<u32> ve_type, ve_ucnt, ve_rpos;
    std::vector<u64> ve_rref;
    // device DRAT stream (proof.cu, proofutils.cuh): one chunk per cacheProof/writeProof pair
    bool proof_en = false;
    std::vector<uint8_t> proofCur;
    std::vector<std::vector<uint8_t>> proofChunks;
    u32 proofCap = 0;

    u32* L(const Clause& c) { return pool.data() + c.off; }
    const u32* L(const Clause& c) const { return pool.data() + c.off; }
};

namespace {

typedef oracle_ctx S;

// ------------------------------------------------------------------ primitives
// primitives.cuh:177-185
void calcSig(S& s, Clause& c) {
    if (c.sz <= 1) return;
    u32 sig = 0;
    const u32* l = s.L(c);
    for (int k = 0; k < c.sz; k++) sig |= MAPHASH(l[k]);
    c.sig = sig;
}

// sclause.cuh:147-168 (membership; clauses are sorted)
bool has(const S& s, const Clause& c, u32 lit) {
    const u32* l = s.L(c);
    for (int k = 0; k < c.sz; k++) if (l[k] == lit) return true;
    return false;
}

// ------------------------------------------------------------------ awaken / prep
// cnf.cu:45-53
void prepCNF(S& s) {
    for (Clause& c : s.cls) {
        std::sort(s.L(c), s.L(c) + c.sz);
        calcSig(s, c);
    }
}

// cnf.cu:33-43 + histogram.cu:54-72 : occurrences of every literal in live clauses
void histSimp(S& s) {
    std::fill(s.hist.begin(), s.hist.end(), 0);
    for (const Clause& c : s.cls) {
        if (c.deleted()) continue;
        const u32* l = s.L(c);
        for (int k = 0; k < c.sz; k++) s.hist[l[k]]++;
    }
}

// occurrence.cu:50-62 ; policy: clause-index order inside a list
void createOT(S& s) {
    for (OL& ol : s.ot) ol.clear();
    for (u32 i = 0; i < s.cls.size(); i++) {
        const Clause& c = s.cls[i];
        if (c.deleted()) continue;
        const u32* l = s.L(c);
        for (int k = 0; k < c.sz; k++) s.ot[l[k]].push_back(i);
    }
}

// count.cu:83-106
void countAll(S& s, u64& nc, u64& nl) {
    nc = nl = 0;
    for (const Clause& c : s.cls)
        if (!c.deleted()) nc++, nl += c.sz;
}

// recycle.cu:60-105 via cnf.cu:129-144 (reallocCNF(true)) : order-preserving compaction,
// storage shrunk to the current sizes, logical capacities reset
void compactCNF(S& s) {
    const u64 maxAddedCls = s.o.ve_en ? s.numClauses : 0;
    const u64 maxAddedLits = s.o.ve_en ? u64(double(s.orgLiterals) * s.o.lits_mul) : 0;
    s.refs_cap = s.numClauses + maxAddedCls;
    s.data_cap = s.refs_cap * NBUCKETS + (s.numLiterals + maxAddedLits);
    std::vector<Clause> ncls;
    std::vector<u32> npool;
    ncls.reserve(s.cls.size());
    npool.reserve(s.pool.size());
    u64 ref = 0;
    for (const Clause& c : s.cls) {
        if (c.deleted()) continue;
        Clause d = c;
        d.off = npool.size();
        d.ref = ref;
        npool.insert(npool.end(), s.L(c), s.L(c) + c.sz);
        ref += NBUCKETS + c.sz;
        ncls.push_back(d);
    }
    s.cls.swap(ncls);
    s.pool.swap(npool);
    s.data_size = ref;
    s.compacted = true;
}

// cnf.cu:146-150
void reallocCNF(S& s) {
    const int times = s.phase + 1;
    if (times > 1 && times != s.o.phases && s.o.shrink_rate > 0 && (times % s.o.shrink_rate) == 0) compactCNF(s);
    else s.compacted = false;
}

// ------------------------------------------------------------------ prop (elimbcp.cu)
u32 bcp_lit_val(const std::vector<u32>& state, u32 lit) {  // elimbcp.cu:33-41
    const u32 st = state[ABS(lit)];
    if (!st) return 0;
    const bool sat = SIGN(lit) ? (st == 2) : (st == 1);
    return sat ? 1 : 2;
}

// elimbcp.cu:144-215 ; returns false on conflict
bool prop(S& s) {
    if (!s.nUnits) return true;
    std::vector<u32> state(s.V + 1, 0), front, next;
    bool conflict = false;
    // bcp_seed_k :43-63
    for (u32 i = 0; i < s.nUnits; i++) {
        const u32 u = s.units[i], v = ABS(u), desired = SIGN(u) ? 2u : 1u;
        if (!state[v]) { state[v] = desired; s.eliminated[v] |= FORCED_MASK; front.push_back(FLIP(u)); }
        else if (state[v] != desired) conflict = true;
    }
    // bcp_propagate_k :65-111 (BFS over falsified literals)
    while (!conflict && !front.empty()) {
        next.clear();
        for (u32 ulit : front) {
            for (u32 ci : s.ot[ulit]) {
                const Clause& c = s.cls[ci];
                if (c.deleted()) continue;
                u32 unit = 0; int nunset = 0; bool sat = false;
                const u32* l = s.L(c);
                for (int k = 0; k < c.sz; k++) {
                    const u32 ve = bcp_lit_val(state, l[k]);
                    if (ve == 1) { sat = true; break; }
                    if (ve == 0) { unit = l[k]; if (++nunset > 1) break; }
                }
                if (sat) continue;
                if (!nunset) conflict = true;
                else if (nunset == 1) {
                    const u32 v = ABS(unit), desired = SIGN(unit) ? 2u : 1u;
                    if (!state[v]) {
                        state[v] = desired; s.eliminated[v] |= FORCED_MASK;
                        s.units.push_back(unit); next.push_back(FLIP(unit));
                    }
                    else if (state[v] != desired) conflict = true;
                }
            }
        }
        front.swap(next);
    }
    if (conflict) { s.cnfstate = UNSAT; return false; }
    // bcp_apply_k :121-142
    for (Clause& c : s.cls) {
        if (c.deleted()) continue;
        u32* l = s.L(c);
        u32 sig = 0; int newsz = 0; bool sat = false;
        for (int k = 0; k < c.sz; k++) {
            const u32 lit = l[k], ve = bcp_lit_val(state, lit);
            if (ve == 1) { sat = true; break; }
            if (ve == 2) continue;
            l[newsz++] = lit; sig |= MAPHASH(lit);
        }
        if (sat) c.st = DELETED;
        else { c.sig = sig; c.sz = newsz; }
    }
    // host enqueue :186-201 (duplicates included)
    // every entry goes on the trail (the reference's host loop, elimbcp.cu:186-201, duplicates included: SURVEY B.11); the
    // count of unassigned variables only moves for a variable assigned here for the first time (the reference decrements
    // per entry in release builds - a latent double count that could only ever produce a spurious SAT; deliberate deviation)
    for (u32 u : s.units) { s.trail.push_back(u); if (s.vstate[ABS(u)] != 2) s.unassigned--; s.vstate[ABS(u)] = 2; }
    countAll(s, s.numClauses, s.numLiterals);
    if (s.numLiterals) histSimp(s);
    createOT(s);
    s.units.clear();
    s.nUnits = 0;
    return true;
}

// ------------------------------------------------------------------ LCVE (lcve.cu)
// lcve.cu:33-62
bool depFreeze(S& s, const OL& ol, u32 cand) {
    const size_t savedTail = s.frozenList.size();
    for (u32 ci : ol) {
        const Clause& c = s.cls[ci];
        if (c.deleted()) continue;
        if (c.sz > s.o.lcve_clause_max) {
            for (size_t k = savedTail; k < s.frozenList.size(); k++) s.frozen[s.frozenList[k]] = 0;
            s.frozenList.resize(savedTail);
            return false;
        }
        const u32* l = s.L(c);
        for (int k = 0; k < c.sz; k++) {
            const u32 v = ABS(l[k]);
            if (!s.frozen[v] && v != cand) { s.frozen[v] = 1; s.frozenList.push_back(v); }
        }
    }
    return true;
}

// lcve.cu:280-398 ; returns false when too few variables were elected
bool LCVE(S& s) {
    // varReorder :280-300 ; assign_scores :221-234 ; GPU_LCV_CMP key.cuh:33-42
    for (u32 v = 1; v <= s.V; v++) {
        s.eligible[v - 1] = v;
        s.scores[v] = s.hist[V2L(v)] * s.hist[NEG(V2L(v))];  // uint32 wrap
    }
    std::sort(s.eligible.begin(), s.eligible.begin() + s.V, [&](u32 a, u32 b) {
        const u32 x = s.scores[a], y = s.scores[b];
        if (x != y) return x < y;
        return a < b;
    });
    const u32 pmax = s.o.mu_pos << s.multiplier, nmax = s.o.mu_neg << s.multiplier;
    const u32 maxoccurs = s.o.lcve_max_occurs;
    s.elected.clear();
    s.frozenList.clear();
    if (s.o.lcve_fast) {
        // -lcvefast: mis_init_k / mis_round_k / mis_freeze_k (lcve.cu:150-217, 338-366).  The two stop conditions and the
        // clause-size limit are FILTERS of the candidate set; rounds of "a candidate wins when no undecided candidate of
        // lower rank shares a clause with it, winners freeze every neighbour" end in the greedy maximal independent set of
        // the rank order.  The set is deterministic; the reference appends winners and frozen variables through atomics
        // (insertAggr, mis_collect_k), so its ORDER is not - here: elected in rank order, frozen list in variable order.
        std::vector<uint8_t> state(s.V + 1, 0);   // 0 none, 1 undecided, 2 elected, 3 frozen
        auto oversize = [&](const OL& ol) {
            for (u32 ci : ol) { const Clause& c = s.cls[ci]; if (!c.deleted() && c.sz > s.o.lcve_clause_max) return true; }
            return false;
        };
        for (u32 v = 1; v <= s.V; v++) {
            if (s.vstate[v] || (!s.assumed.empty() && s.assumed[v])) continue;
            const u32 p = V2L(v), n = NEG(p), ps = s.hist[p], ns = s.hist[n];
            if ((ps || ns) && ps <= maxoccurs && ns <= maxoccurs && !(ps >= pmax && ns >= nmax) && !oversize(s.ot[p]) && !oversize(s.ot[n]))
                state[v] = 1;
        }
        for (u32 ei = 0; ei < s.V; ei++) {
            const u32 cand = s.eligible[ei];
            if (state[cand] != 1) continue;
            state[cand] = 2;
            s.elected.push_back(cand);
            for (int side = 0; side < 2; side++)
                for (u32 ci : s.ot[V2L(cand) | (u32)side]) {
                    const Clause& c = s.cls[ci];
                    if (c.deleted()) continue;
                    const u32* l = s.L(c);
                    for (int k = 0; k < c.sz; k++) {
                        const u32 u = ABS(l[k]);
                        if (u == cand) continue;
                        s.frozen[u] = 1;
                        if (state[u] == 1) state[u] = 3;
                    }
                }
        }
        for (u32 v = 1; v <= s.V; v++) if (s.frozen[v]) s.frozenList.push_back(v);
    }
    else
    // lcve_k :64-101
    for (u32 ei = 0; ei < s.V; ei++) {
        const u32 cand = s.eligible[ei];
        if (s.frozen[cand]) continue;
        if (s.vstate[cand]) continue;
        if (!s.assumed.empty() && s.assumed[cand]) continue;   // lcve.cu:88
        const u32 p = V2L(cand), n = NEG(p);
        const u32 ps = s.hist[p], ns = s.hist[n];
        if (!ps && !ns) continue;
        if (ps > maxoccurs || ns > maxoccurs) break;
        if (ps >= pmax && ns >= nmax) break;
        if (depFreeze(s, s.ot[p], cand) && depFreeze(s, s.ot[n], cand)) s.elected.push_back(cand);
    }
    s.numElected = u32(s.elected.size());
    if (s.keep_snaps) s.electedLog.push_back(s.elected);
    // mapFrozen :400-419 ; varcore aliases eligible (simplify.cu:118)
    if (s.o.ve_fun_en && !s.varcore_dead) {
        if (s.frozenList.empty()) { s.varcore = nullptr; s.varcore_dead = true; }
        else {
            s.varcore = s.eligible.data();
            for (u32 i = 0; i < s.frozenList.size(); i++) s.eligible[s.frozenList[i]] = i;
        }
    }
    // clearFrozen solver.hpp:291-297
    for (u32 v : s.frozenList) s.frozen[v] = 0;
    return s.numElected >= s.o.lcve_min_vars;
}

// segsort.cu:37-48 ; OLIST_CMP key.cuh:67-83.  The call passes d_segs+3 as the segment *starts*, so
// moderngpu's implicit first segment [0, segs[3]) is list 2 (lists 0 and 1 are empty): every list is sorted.
void sortOT(S& s) {
    for (u32 lit = 2; lit < s.ot.size(); lit++) {
        OL& ol = s.ot[lit];
        if (ol.size() < 2) continue;
        std::sort(ol.begin(), ol.end(), [&](u32 a, u32 b) {
            const Clause& x = s.cls[a];
            const Clause& y = s.cls[b];
            if (x.sz != y.sz) return u32(x.sz) < u32(y.sz);
            const u32 x0 = s.L(x)[0], y0 = s.L(y)[0];
            if (x0 != y0) return x0 < y0;
            const u32 xb = s.L(x)[x.sz - 1], yb = s.L(y)[y.sz - 1];
            if (xb != yb) return xb < yb;
            if (x.sig != y.sig) return x.sig < y.sig;
            return x.ref < y.ref;
        });
    }
}

// ------------------------------------------------------------------ merges (elimination.cuh)
// elimination.cuh:180-205 : resolvent length, 0 if tautology
int merge_len(const S& s, u32 x, const Clause& c1, const Clause& c2) {
    const int n1 = c1.sz, n2 = c2.sz;
    const u32* a = s.L(c1); const u32* b = s.L(c2);
    int it1 = 0, it2 = 0, len = n1 + n2 - 2;
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = ABS(lit1), v2 = ABS(lit2);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (IS_TAUTOLOGY(lit1, lit2)) return 0;
        else if (v1 < v2) it1++;
        else if (v2 < v1) it2++;
        else { it1++, it2++; len--; }
    }
    return len;
}

// elimination.cuh:277-310 : write the resolvent, return its length (0 if tautology)
int merge_out(const S& s, u32 x, const u32* a, int n1, const u32* b, int n2, u32* out) {
    int it1 = 0, it2 = 0, len = 0;
    while (it1 < n1 && it2 < n2) {
        const u32 lit1 = a[it1], lit2 = b[it2], v1 = ABS(lit1), v2 = ABS(lit2);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (IS_TAUTOLOGY(lit1, lit2)) return 0;
        else if (v1 < v2) { it1++; out[len++] = lit1; }
        else if (v2 < v1) { it2++; out[len++] = lit2; }
        else { it1++, it2++; out[len++] = lit1; }
    }
    while (it1 < n1) { const u32 l = a[it1++]; if (ABS(l) != x) out[len++] = l; }
    while (it2 < n2) { const u32 l = b[it2++]; if (ABS(l) != x) out[len++] = l; }
    (void)s;
    return len;
}

// elimination.cuh:109-131
bool isTautology(const S& s, u32 x, const Clause& c1, const Clause& c2) {
    const u32* a = s.L(c1); const u32* b = s.L(c2);
    int it1 = 0, it2 = 0;
    while (it1 < c1.sz && it2 < c2.sz) {
        const u32 v1 = ABS(a[it1]), v2 = ABS(b[it2]);
        if (v1 == x) it1++;
        else if (v2 == x) it2++;
        else if (IS_TAUTOLOGY(a[it1], b[it2])) return true;
        else if (v1 < v2) it1++;
        else if (v2 < v1) it2++;
        else { it1++; it2++; }
    }
    return false;
}

void countLitsBefore(const S& s, const OL& list, u32& n) {  // elimination.cuh:356-363
    for (u32 ci : list) if (s.cls[ci].original()) n += s.cls[ci].sz;
}

// ------------------------------------------------------------------ witness stack (model.cuh)
void saveWitness(S& s, u32 witness) {  // model.cuh:29-34
    s.resolved.push_back(V2L(s.vorg[ABS(witness)]) | SIGN(witness));
    s.resolved.push_back(1);
}
void saveClause(S& s, const Clause& c, u32 witlit) {  // model.cuh:36-53
    const size_t first = s.resolved.size();
    size_t wpos = first;
    const u32* l = s.L(c);
    for (int k = 0; k < c.sz; k++) {
        if (l[k] == witlit) wpos = s.resolved.size();
        s.resolved.push_back(V2L(s.vorg[ABS(l[k])]) | SIGN(l[k]));
    }
    std::swap(s.resolved[first], s.resolved[wpos]);
    s.resolved.push_back(u32(c.sz));
}

// elimination.cuh:443-492 : save the smaller side, delete everything, clear the lists
void toblivion_save(S& s, u32 p, u32 n, u32 pOrgs, u32 nOrgs, OL& poss, OL& negs) {
    const bool which = pOrgs > nOrgs;
    if (which) {
        for (u32 ci : negs) { Clause& c = s.cls[ci]; if (c.original()) saveClause(s, c, n); c.st = DELETED; }
        saveWitness(s, p);
    } else {
        for (u32 ci : poss) { Clause& c = s.cls[ci]; if (c.original()) saveClause(s, c, p); c.st = DELETED; }
        saveWitness(s, n);
    }
    OL& other = which ? poss : negs;
    for (u32 ci : other) s.cls[ci].st = DELETED;
    poss.clear(); negs.clear();
}
// elimination.cuh:494-503
void toblivion(S& s, OL& poss, OL& negs) {
    for (u32 ci : poss) s.cls[ci].st = DELETED;
    for (u32 ci : negs) s.cls[ci].st = DELETED;
    poss.clear(); negs.clear();
}
// elimination.cuh:505-550 (which = list sizes)  and :552-594 (which = pOrgs > nOrgs)
void saveResolved(S& s, u32 p, u32 n, bool which, const OL& poss, const OL& negs) {
    if (which) {
        for (u32 ci : negs) { const Clause& c = s.cls[ci]; if (c.original()) saveClause(s, c, n); }
        saveWitness(s, p);
    } else {
        for (u32 ci : poss) { const Clause& c = s.cls[ci]; if (c.original()) saveClause(s, c, p); }
        saveWitness(s, n);
    }
}

void freezeBinaries(S& s, const OL& list) {  // elimination.cuh:73-79
    for (u32 ci : list) { Clause& c = s.cls[ci]; if (c.original() && c.sz == 2) c.molten = 0; }
}
void freezeClauses(S& s, const OL& poss, const OL& negs) {  // elimination.cuh:81-93
    for (u32 ci : poss) { Clause& c = s.cls[ci]; if (c.original() && c.molten) c.molten = 0; }
    for (u32 ci : negs) { Clause& c = s.cls[ci]; if (c.original() && c.molten) c.molten = 0; }
}

// ------------------------------------------------------------------ device DRAT stream (proofutils.cuh)
// bytes of the 7-bit variable-length encoding of the ORIGINAL literal (proof.cu:31-41 BLUT, :57-63)
u32 proofLitBytes(const S& s, u32 lit) {
    u32 org = V2L(s.vorg[ABS(lit)]) | SIGN(lit);
    u32 n = 1;
    while (org & 0xFFFFFF80u) { n++; org >>= 7; }
    return n;
}
// proofutils.cuh:96-121
void saveProofLiteral(S& s, u32 lit) {
    u32 org = V2L(s.vorg[ABS(lit)]) | SIGN(lit);
    while (org & 0xFFFFFF80u) { s.proofCur.push_back(uint8_t((org & 0x7Fu) | 0x80u)); org >>= 7; }
    s.proofCur.push_back(uint8_t(org));
}
// proofutils.cuh:123-170
void saveProofClause(S& s, const u32* lits, int n, uint8_t state) {
    s.proofCur.push_back(state);
    for (int k = 0; k < n; k++) saveProofLiteral(s, lits[k]);
    s.proofCur.push_back(0);
}
void saveProofClause(S& s, const Clause& c, uint8_t state) { saveProofClause(s, s.L(c), c.sz, state); }
// mergeProof (elimination.cuh:162-215): proof bytes of the resolvent of c1 and c2 on x (prefix + literals + suffix)
u32 mergeProofBytes(const S& s, u32 x, const Clause& c1, const Clause& c2) {
    std::vector<u32> out(size_t(c1.sz) + c2.sz);
    const int n = merge_out(s, x, s.L(c1), c1.sz, s.L(c2), c2.sz, out.data());
    if (!n) return 0;
    u32 bytes = 2;
    for (int k = 0; k < n; k++) bytes += proofLitBytes(s, out[k]);
    return bytes;
}
// the proof guard of the four counting functions (resolve.cuh:66-70, :152-154, elimination.cuh:445-447, function.cuh:232-235)
bool proofGuard(const S& s, u32 nElements, u32 proofBytes) {
    return s.proof_en && (nElements > ADDEDCLS_MAX || proofBytes > ADDEDPROOF_MAX);
}
// cacheProof + writeProof (proof.cu:160-199, 232-247): the device stream leaves as one chunk
void flushProof(S& s) {
    if (!s.proof_en) return;
    s.proofChunks.push_back(s.proofCur);
    s.proofCur.clear();
}

// ------------------------------------------------------------------ resolvent counting
// resolve.cuh:28-109 (no clause bound) ; returns true if resolvable
bool countResolvents_simple(S& s, u32 x, const OL& me, const OL& other, u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    const int rlimit = int(s.o.ve_clause_max);
    u32 proofBytes = 0;
    for (u32 i : me) {
        const Clause& ci = s.cls[i];
        if (ci.learnt()) continue;
        for (u32 j : other) {
            const Clause& cj = s.cls[j];
            if (cj.learnt()) continue;
            const int rsize = merge_len(s, x, ci, cj);
            if (rsize && s.proof_en) proofBytes += mergeProofBytes(s, x, ci, cj);
            if (rsize == 1) nElements++;
            else if (rsize) {
                if (rlimit && rsize > rlimit) return false;
                ++nAddedCls;
                nAddedLits += rsize;
            }
        }
    }
    if (proofGuard(s, nElements, proofBytes)) return false;
    if (nAddedCls > ADDEDCLS_MAX || nAddedLits > ADDEDLITS_MAX) return false;
    return true;
}
// resolve.cuh:111-199 (bounded)
bool countResolvents(S& s, u32 x, u32 nClsBefore, const OL& me, const OL& other, u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    nElements = 0, nAddedCls = 0, nAddedLits = 0;
    const int rlimit = int(s.o.ve_clause_max);
    u32 proofBytes = 0;
    for (u32 i : me) {
        const Clause& ci = s.cls[i];
        if (ci.learnt()) continue;
        for (u32 j : other) {
            const Clause& cj = s.cls[j];
            if (cj.learnt()) continue;
            const int rsize = merge_len(s, x, ci, cj);
            if (rsize && s.proof_en) proofBytes += mergeProofBytes(s, x, ci, cj);
            if (rsize == 1) nElements++;
            else if (rsize) {
                if (++nAddedCls > nClsBefore || (rlimit && rsize > rlimit)) return false;
                nAddedLits += rsize;
            }
        }
    }
    if (proofGuard(s, nElements, proofBytes)) return false;
    if (nAddedCls > ADDEDCLS_MAX || nAddedLits > ADDEDLITS_MAX) return false;
    if (s.o.ve_lbound_en) {
        u32 nLitsBefore = 0;
        countLitsBefore(s, me, nLitsBefore);
        countLitsBefore(s, other, nLitsBefore);
        if (nAddedLits > nLitsBefore) return false;
    }
    return true;
}
// elimination.cuh:365-441 ; returns TRUE when substitution is NOT possible
bool countSubstituted(S& s, u32 x, u32 nClsBefore, const OL& me, const OL& other, u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    const int rlimit = int(s.o.ve_clause_max);
    u32 proofBytes = 0;
    for (u32 i : me) {
        const Clause& ci = s.cls[i];
        if (ci.learnt()) continue;
        const bool ci_m = ci.molten;
        for (u32 j : other) {
            const Clause& cj = s.cls[j];
            if (cj.original() && ci_m != bool(cj.molten)) {
                const int rsize = merge_len(s, x, ci, cj);
                if (rsize && s.proof_en) proofBytes += mergeProofBytes(s, x, ci, cj);
                if (rsize == 1) nElements++;
                else if (rsize) {
                    if (++nAddedCls > nClsBefore || (rlimit && rsize > rlimit)) return true;
                    nAddedLits += rsize;
                }
            }
        }
    }
    if (proofGuard(s, nElements, proofBytes)) return true;
    if (nAddedCls > ADDEDCLS_MAX || nAddedLits > ADDEDLITS_MAX) return true;
    if (s.o.ve_lbound_en) {
        u32 nLitsBefore = 0;
        countLitsBefore(s, me, nLitsBefore);
        countLitsBefore(s, other, nLitsBefore);
        if (nAddedLits > nLitsBefore) return true;
    }
    return false;
}
// function.cuh:181-257 ; TRUE when not possible
bool countCoreSubstituted(S& s, u32 x, u32 nClsBefore, const OL& me, const OL& other, u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    const int rlimit = int(s.o.ve_clause_max);
    u32 proofBytes = 0;
    for (u32 i : me) {
        const Clause& ci = s.cls[i];
        if (ci.learnt()) continue;
        const bool ci_m = ci.molten;
        for (u32 j : other) {
            const Clause& cj = s.cls[j];
            if (cj.original() && (!ci_m || !cj.molten)) {
                const int rsize = merge_len(s, x, ci, cj);
                if (rsize && s.proof_en) proofBytes += mergeProofBytes(s, x, ci, cj);
                if (rsize == 1) nElements++;
                else if (rsize) {
                    if (++nAddedCls > nClsBefore || (rlimit && rsize > rlimit)) return true;
                    nAddedLits += rsize;
                }
            }
        }
    }
    if (proofGuard(s, nElements, proofBytes)) return true;
    if (nAddedCls > ADDEDCLS_MAX || nAddedLits > ADDEDLITS_MAX) return true;
    if (s.o.ve_lbound_en) {
        u32 nLitsBefore = 0;
        countLitsBefore(s, me, nLitsBefore);
        countLitsBefore(s, other, nLitsBefore);
        if (nAddedLits > nLitsBefore) return true;
    }
    return false;
}

// ------------------------------------------------------------------ gates
// equivalence.cuh:111-130
u32 find_sfanin(S& s, u32 gate_out, const OL& list) {
    u32 imp = 0; int nImps = 0;
    for (u32 ci : list) {
        Clause& c = s.cls[ci];
        if (c.original() && c.sz == 2) {
            const u32* l = s.L(c);
            imp = FLIP(l[0] ^ l[1] ^ gate_out);
            c.molten = 1;
            nImps++;
        }
        if (nImps > 1) return 0;
    }
    return imp;
}
// equivalence.cuh:132-171
u32 find_equ_gate(S& s, u32 p, u32 n, const OL& poss, const OL& negs) {
    if (s.cls[poss[0]].sz > 2 || s.cls[negs[0]].sz > 2) return 0;
    u32 first = find_sfanin(s, p, poss);
    if (first) {
        u32 second = n; const u32 def = first;
        if (second < first) first = second, second = def;
        for (u32 ci : negs) {
            Clause& c = s.cls[ci];
            const u32* l = s.L(c);
            if (c.original() && c.sz == 2 && l[0] == first && l[1] == second) { c.molten = 1; return def; }
        }
    }
    freezeBinaries(s, poss);
    return 0;
}
// equivalence.cuh:28-57
void substitute_single_clause(S& s, u32 dx, u32 def, Clause& org, u32& nUnits) {
    u32* l = s.L(org);
    int n = 0;
    for (int k = 0; k < org.sz; k++) {
        const u32 lit = l[k];
        if (lit == dx) l[n++] = def;
        else if (lit != def) l[n++] = lit;
    }
    org.sz = n;
    if (n == 1) nUnits++;
    else { std::sort(l, l + n); calcSig(s, org); }
}
void appendUnits(S& s, const OL& ol) {  // elimination.cuh:95-105
    for (u32 ci : ol) { const Clause& c = s.cls[ci]; if (c.sz == 1) s.units.push_back(s.L(c)[0]); }
}
// equivalence.cuh:59-109
void substitute_single(S& s, u32 p, u32 n, u32 def, const OL& poss, const OL& negs) {
    const u32 def_f = FLIP(def);
    u32 nNegUnits = 0, nPosUnits = 0;
    for (u32 ci : negs) {
        Clause& c = s.cls[ci];
        if (c.learnt() || c.molten || has(s, c, def)) c.st = DELETED;
        else substitute_single_clause(s, n, def_f, c, nNegUnits);
    }
    for (u32 ci : poss) {
        Clause& c = s.cls[ci];
        if (c.learnt() || c.molten || has(s, c, def_f)) c.st = DELETED;
        else substitute_single_clause(s, p, def, c, nPosUnits);
    }
    if (nPosUnits || nNegUnits) {
        if (nNegUnits) appendUnits(s, negs);
        if (nPosUnits) appendUnits(s, poss);
    }
    if (s.proof_en) {   // equivalence.cuh:100-109, addProof proofutils.cuh:172-180
        for (u32 ci : negs) if (s.cls[ci].original()) saveProofClause(s, s.cls[ci], PROOF_ADDED);
        for (u32 ci : poss) if (s.cls[ci].original()) saveProofClause(s, s.cls[ci], PROOF_ADDED);
    }
}

// and.cuh:28-44
int find_fanin(S& s, u32 gate_out, const OL& list, std::vector<u32>& out_c, u32& sig) {
    sig = 0; int nImps = 0;
    for (u32 ci : list) {
        Clause& c = s.cls[ci];
        if (c.original() && c.sz == 2) {
            const u32* l = s.L(c);
            const u32 imp = FLIP(l[0] ^ l[1] ^ gate_out);
            out_c[nImps++] = imp;
            sig |= MAPHASH(imp);
            c.molten = 1;
        }
    }
    return nImps;
}
// and.cuh:46-113
bool find_ao_gate(S& s, u32 dx, const OL& dx_list, u32 fx, const OL& fx_list, u32 nOrgCls,
                  std::vector<u32>& out_c, u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    if (s.cls[dx_list[0]].sz > 2 || s.cls[fx_list.back()].sz < 3) return false;
    u32 sig;
    if (out_c.size() < dx_list.size() + 2) out_c.resize(dx_list.size() + 2);
    int nImps = find_fanin(s, dx, dx_list, out_c, sig);
    if (nImps > 1) {
        const u32 x = ABS(dx);
        out_c[nImps++] = fx;
        sig |= MAPHASH(fx);
        std::sort(out_c.begin(), out_c.begin() + nImps);
        for (u32 ci : fx_list) {
            Clause& c = s.cls[ci];
            if (c.original() && c.sz == nImps && SUBSIG(c.sig, sig) && std::equal(s.L(c), s.L(c) + nImps, out_c.begin())) {
                c.molten = 1;
                nElements = 0, nAddedCls = 0, nAddedLits = 0;
                if (countSubstituted(s, x, nOrgCls, dx_list, fx_list, nElements, nAddedCls, nAddedLits)) { c.molten = 0; break; }
                return true;
            }
        }
    }
    freezeBinaries(s, dx_list);
    return false;
}

// ifthenelse.cuh:28-50
i64 fast_equality_check(S& s, u32 x, u32 y, u32 z) {
    if (s.ot[y].size() > s.ot[z].size()) std::swap(y, z);
    if (s.ot[x].size() > s.ot[y].size()) std::swap(x, y);
    const OL& list = s.ot[x];
    u32 t[3] = {x, y, z};
    std::sort(t, t + 3);
    for (u32 ci : list) {
        const Clause& c = s.cls[ci];
        if (c.molten) continue;
        const u32* l = s.L(c);
        if (c.original() && c.sz == 3 && l[0] == t[0] && l[1] == t[1] && l[2] == t[2]) return ci;
    }
    return -1;
}
// ifthenelse.cuh:52-125
bool find_ite_gate(S& s, u32 dx, const OL& dx_list, u32 fx, const OL& fx_list, u32 nOrgCls,
                   u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    if (s.cls[dx_list.back()].sz == 2) return false;
    const u32 v = ABS(dx);
    for (size_t i = 0; i < dx_list.size(); i++) {
        Clause& ci = s.cls[dx_list[i]];
        if (!(ci.original() && ci.sz == 3)) continue;
        u32 xi = s.L(ci)[0], yi = s.L(ci)[1], zi = s.L(ci)[2];
        if (yi == dx) std::swap(xi, yi);
        if (zi == dx) std::swap(xi, zi);
        for (size_t j = i + 1; j < dx_list.size(); j++) {
            Clause& cj = s.cls[dx_list[j]];
            if (!(cj.original() && cj.sz == 3)) continue;
            u32 xj = s.L(cj)[0], yj = s.L(cj)[1], zj = s.L(cj)[2];
            if (yj == dx) std::swap(xj, yj);
            if (zj == dx) std::swap(xj, zj);
            if (ABS(yi) == ABS(zj)) std::swap(yj, zj);
            if (ABS(zi) == ABS(zj)) continue;
            if (yi != FLIP(yj)) continue;
            const i64 r1 = fast_equality_check(s, fx, yi, FLIP(zi));
            if (r1 < 0) continue;
            const i64 r2 = fast_equality_check(s, fx, yj, FLIP(zj));
            if (r2 < 0) continue;
            ci.molten = 1, cj.molten = 1;
            s.cls[r1].molten = 1, s.cls[r2].molten = 1;
            nElements = 0, nAddedCls = 0, nAddedLits = 0;
            if (countSubstituted(s, v, nOrgCls, dx_list, fx_list, nElements, nAddedCls, nAddedLits)) {
                ci.molten = 0, cj.molten = 0;
                s.cls[r1].molten = 0, s.cls[r2].molten = 0;
                return false;
            }
            return true;
        }
    }
    return false;
}

// xor.cuh:37-56
void freeze_arities(S& s, const OL& me, const OL& other) {
    for (u32 ci : me) { Clause& c = s.cls[ci]; if (c.sz > 2 && c.molten) c.molten = 0; }
    for (u32 ci : other) { Clause& c = s.cls[ci]; if (c.sz > 2 && c.molten) c.molten = 0; }
}
// xor.cuh:58-74
bool checkArity(const S& s, const Clause& c, const u32* literals, int size) {
    const u32* l = s.L(c);
    for (int k = 0; k < c.sz; k++) {
        int j = 0;
        for (; j < size; j++) if (l[k] == literals[j]) break;
        if (j == size) return false;
    }
    return true;
}
// xor.cuh:76-109
bool makeArity(S& s, u32& parity, u32* literals, int size) {
    const u32 oldparity = parity;
    while (__builtin_popcount(++parity) & 1) {}
    for (int k = 0; k < size; k++) {
        const u32 bit = k < 32 ? (1u << k) : 0u;   // xor.cuh:78 `(1UL << k)` truncated to 32 bits: 0 from bit 32 on
        if ((parity & bit) != (oldparity & bit)) literals[k] = FLIP(literals[k]);
    }
    u32 best = literals[0];
    int minsize = int(s.ot[best].size());
    for (int k = 1; k < size; k++) {
        const int lsize = int(s.ot[literals[k]].size());
        if (lsize < minsize) { minsize = lsize; best = literals[k]; }
    }
    for (u32 ci : s.ot[best]) {
        Clause& c = s.cls[ci];
        if (c.original() && c.sz == size && checkArity(s, c, literals, size)) { c.molten = 1; return true; }
    }
    return false;
}
// xor.cuh:111-185
bool find_xor_gate(S& s, u32 dx, const OL& dx_list, u32 fx, const OL& fx_list, u32 nOrgCls,
                   u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    (void)fx;
    if (s.cls[dx_list.back()].sz == 2 || s.cls[fx_list.back()].sz == 2) return false;
    const int maxarity = int(s.o.xor_max_arity);
    if (s.cls[dx_list[0]].sz - 1 > maxarity) return false;
    const u32 v = ABS(dx);
    std::vector<u32> out_c;
    for (u32 idx : dx_list) {
        Clause& ci = s.cls[idx];
        if (!ci.original()) continue;
        const int size = ci.sz, arity = size - 1;
        if (size < 3 || arity > maxarity) continue;
        out_c.assign(s.L(ci), s.L(ci) + size);
        u32 parity = 0;
        // xor.cuh:148 `1 << arity` runs on the GPU: a 32-bit shift by 32 or more gives 0 there (PTX shl clamps the
        // amount), not the x86 result; only reachable with --xormaxarity > 31
        int itargets = arity >= 32 ? 0 : int(1u << arity);
        while (--itargets && makeArity(s, parity, out_c.data(), size)) {}
        if (itargets) freeze_arities(s, dx_list, fx_list);
        else {
            ci.molten = 1;
            nElements = 0, nAddedCls = 0, nAddedLits = 0;
            if (countSubstituted(s, v, nOrgCls, dx_list, fx_list, nElements, nAddedCls, nAddedLits)) {
                freeze_arities(s, dx_list, fx_list);
                break;
            }
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ function tables (function.cuh)
const int MAXFUNVAR = 12;
const int FUNTABLEN = 64;
typedef u64 Fun[FUNTABLEN];
const u64 MAGIC[6] = {0xaaaaaaaaaaaaaaaaULL, 0xccccccccccccccccULL, 0xf0f0f0f0f0f0f0f0ULL,
                      0xff00ff00ff00ff00ULL, 0xffff0000ffff0000ULL, 0xffffffff00000000ULL};
void fillfun(Fun f, u64 v) { for (int i = 0; i < FUNTABLEN; i++) f[i] = v; }
bool isfalsefun(const Fun f) { for (int i = 0; i < FUNTABLEN; i++) if (f[i]) return false; return true; }
// function.cuh:86-112
void clause2fun(int v, bool sign, Fun f) {
    if (v < 6) {
        u64 val = MAGIC[v];
        if (sign) val = ~val;
        for (int i = 0; i < FUNTABLEN; i++) f[i] |= val;
    } else {
        u64 val = sign ? ~0ULL : 0ULL;
        int j = 0;
        const int sv = 1 << (v - 6);
        for (int i = 0; i < FUNTABLEN; i++) {
            f[i] |= val;
            if (++j >= sv) { val = ~val; j = 0; }
        }
    }
}
// function.cuh:114-148
bool buildfuntab_all(S& s, u32 lit, Fun f) {
    Fun cls;
    fillfun(f, ~0ULL);
    for (u32 ci : s.ot[lit]) {
        const Clause& c = s.cls[ci];
        if (c.learnt()) continue;
        fillfun(cls, 0);
        const u32* l = s.L(c);
        for (int k = 0; k < c.sz; k++) {
            const u32 other = l[k];
            if (other == lit) continue;
            const u32 mvar = s.varcore[ABS(other)];
            if (mvar >= u32(MAXFUNVAR)) return false;
            clause2fun(int(mvar), SIGN(other), cls);
        }
        for (int i = 0; i < FUNTABLEN; i++) f[i] &= cls[i];
    }
    return true;
}
// function.cuh:150-179
void buildfuntab_tail(S& s, u32 lit, int tail, const OL& ol, Fun fun, bool& core) {
    Fun cls;
    for (int j = 0; j < tail; ++j) {
        const Clause& c = s.cls[ol[j]];
        if (c.learnt()) continue;
        fillfun(cls, 0);
        const u32* l = s.L(c);
        for (int k = 0; k < c.sz; k++) {
            const u32 other = l[k];
            if (other == lit) continue;
            const u32 mvar = s.varcore[ABS(other)];
            clause2fun(int(mvar), SIGN(other), cls);
        }
        for (int i = 0; i < FUNTABLEN; i++) fun[i] &= cls[i];
    }
    if (isfalsefun(fun)) { s.cls[ol[tail]].molten = 1; core = true; }
}
// function.cuh:275-327
bool find_fun_gate(S& s, u32 p, u32 n, u32 nOrgCls, u32& nElements, u32& nAddedCls, u32& nAddedLits) {
    Fun pos, neg;
    if (buildfuntab_all(s, p, pos) && buildfuntab_all(s, n, neg)) {
        u64 allzero = 0;
        for (int i = 0; i < FUNTABLEN; i++) allzero |= (pos[i] & neg[i]);
        if (!allzero) {
            u64* fun = pos;
            const OL& poss = s.ot[p];
            bool core = false;
            for (int i = int(poss.size()) - 1; i >= 0; i--) {
                memcpy(fun, neg, sizeof(Fun));
                if (s.cls[poss[i]].original()) buildfuntab_tail(s, p, i, poss, fun, core);
            }
            const OL& negs = s.ot[n];
            for (int i = int(negs.size()) - 1; i >= 0; i--) {
                fillfun(fun, ~0ULL);
                if (s.cls[negs[i]].original()) buildfuntab_tail(s, n, i, negs, fun, core);
            }
            nElements = 0, nAddedCls = 0, nAddedLits = 0;
            if (countCoreSubstituted(s, ABS(p), nOrgCls, poss, negs, nElements, nAddedCls, nAddedLits)) {
                if (core) freezeClauses(s, poss, negs);
                return false;
            }
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ BVE phase 1 (bounded.cuh:282-394)
inline u32 ENCODEVARINFO(u32 t, u32 cls, u32 lits) { return t | (cls << 2) | (lits << 16); }

void variable_elimination(S& s, u32 tid, u32 x, u32 pOrgs, u32 nOrgs) {
    const u32 p = V2L(x), n = NEG(p);
    OL& poss = s.ot[p]; OL& negs = s.ot[n];
    u32 elimType = 0, nElements = 0, nAddedCls = 0, nAddedLits = 0;
    auto none = [&]() { s.ve_ucnt[tid] = 0, s.ve_type[tid] = 0, s.ve_rref[tid] = 0, s.ve_rpos[tid] = 0; };
    if (!pOrgs || !nOrgs) {
        toblivion_save(s, p, n, pOrgs, nOrgs, poss, negs);
        none(); s.eliminated[x] |= MELTING_MASK;
        return;
    }
    if (u32 def = find_equ_gate(s, p, n, poss, negs)) {
        saveResolved(s, p, n, pOrgs > nOrgs, poss, negs);
        substitute_single(s, p, n, def, poss, negs);
        none(); s.eliminated[x] |= MELTING_MASK;
        return;
    }
    if ((pOrgs == 1 || nOrgs == 1) && countResolvents_simple(s, x, poss, negs, nElements, nAddedCls, nAddedLits)) {
        if (nAddedCls) {
            s.ve_type[tid] = ENCODEVARINFO(RES_MASK, nAddedCls, nAddedLits);
            s.ve_ucnt[tid] = nElements, s.ve_rpos[tid] = nAddedCls, s.ve_rref[tid] = nAddedLits + NBUCKETS * nAddedCls;
        } else {
            toblivion_save(s, p, n, pOrgs, nOrgs, poss, negs);
            none(); s.eliminated[x] |= MELTING_MASK;
        }
        return;
    }
    const u32 nClsBefore = pOrgs + nOrgs;
    elimType = 0, nElements = 0, nAddedCls = 0, nAddedLits = 0;
    std::vector<u32> outs;
    if (nClsBefore > 2) {
        if (nOrgs < s.o.sh_max_bve_out1 && find_ao_gate(s, n, negs, p, poss, nClsBefore, outs, nElements, nAddedCls, nAddedLits))
            elimType = AOIX_MASK;
        else if (!nAddedCls && pOrgs < s.o.sh_max_bve_out1 && find_ao_gate(s, p, poss, n, negs, nClsBefore, outs, nElements, nAddedCls, nAddedLits))
            elimType = AOIX_MASK;
    }
    if (!elimType && nClsBefore > 3) {
        if (find_ite_gate(s, p, poss, n, negs, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
        else if (!nAddedCls && find_ite_gate(s, n, negs, p, poss, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
        else if (find_xor_gate(s, p, poss, n, negs, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
        else if (!nAddedCls && find_xor_gate(s, n, negs, p, poss, nClsBefore, nElements, nAddedCls, nAddedLits)) elimType = AOIX_MASK;
    }
    if (s.varcore && !elimType && nClsBefore > 2 && find_fun_gate(s, p, n, nClsBefore, nElements, nAddedCls, nAddedLits))
        elimType = CORE_MASK;
    else if (!elimType && !nAddedCls && countResolvents(s, x, nClsBefore, poss, negs, nElements, nAddedCls, nAddedLits))
        elimType = RES_MASK;
    if (!nAddedCls) {
        toblivion_save(s, p, n, pOrgs, nOrgs, poss, negs);
        none(); s.eliminated[x] |= MELTING_MASK;
    } else if (elimType) {
        s.ve_type[tid] = ENCODEVARINFO(elimType, nAddedCls, nAddedLits);
        s.ve_ucnt[tid] = nElements, s.ve_rpos[tid] = nAddedCls, s.ve_rref[tid] = nAddedLits + NBUCKETS * nAddedCls;
    } else none();
}

// bounded.cuh:171-280 (resolve_x / substitute_x / coresubstitute_x share one loop shape)
void emit_resolvents(S& s, u32 x, u32 elimType, u32 nAddedCls, u32 addedPos, u64 addedRef, OL& me, OL& other) {
    const u32 checksum = addedPos + nAddedCls;
    u64 newref = addedRef;
    std::vector<u32> out;
    for (size_t i = 0; i < me.size() && addedPos < checksum; i++) {
        const Clause ci = s.cls[me[i]];
        if (ci.learnt()) continue;
        for (size_t j = 0; j < other.size() && addedPos < checksum; j++) {
            const Clause cj = s.cls[other[j]];
            if (cj.learnt()) continue;
            if (elimType == AOIX_MASK && bool(ci.molten) == bool(cj.molten)) continue;
            if (elimType == CORE_MASK && ci.molten && cj.molten) continue;
            out.resize(size_t(ci.sz) + cj.sz);
            const int rsize = merge_out(s, x, s.L(ci), ci.sz, s.L(cj), cj.sz, out.data());
            if (!rsize) continue;
            if (rsize == 1) {   // bounded.cuh:76-81, 99-105: unit + saveProofUnit
                s.units.push_back(out[0]);
                if (s.proof_en) saveProofClause(s, out.data(), 1, PROOF_ADDED);
                continue;
            }
            if (s.proof_en) saveProofClause(s, out.data(), rsize, PROOF_ADDED);   // bounded.cuh:91-92, 117-118
            // new SCLAUSE: ORIGINAL, added, sorted, sig (bounded.cuh:86-120)
            Clause a;
            a.st = ORIGINAL, a.molten = 0, a.added = 1, a.usage = 0, a.lbd = 0, a.sig = 0, a.sz = rsize;
            a.off = s.pool.size();
            a.ref = newref;
            s.pool.insert(s.pool.end(), out.begin(), out.begin() + rsize);
            for (int k = 0; k < rsize; k++) a.sig |= MAPHASH(out[k]);
            // refs[addedPos] = newref : the scans make positions consecutive over elected order
            if (s.cls.size() != addedPos) { fprintf(stderr, "oracle: resolvent position mismatch (%zu vs %u)\n", s.cls.size(), addedPos); abort(); }
            s.cls.push_back(a);
            addedPos++;
            newref += rsize + NBUCKETS;
        }
    }
    toblivion(s, me, other);
}

// elimination.cu:235-266 + bounded.cuh:397-544
void VE(S& s) {
    const u32 E = s.numElected;
    s.ve_type.assign(E, 0), s.ve_ucnt.assign(E, 0), s.ve_rpos.assign(E, 0), s.ve_rref.assign(E, 0);
    const bool in = s.o.sigma_calls > 1;
    // phase 1
    for (u32 tid = 0; tid < E; tid++) {
        const u32 x = s.elected[tid], p = V2L(x), n = NEG(p);
        u32 pOrgs = 0, nOrgs = 0;
        if (in) {
            for (u32 ci : s.ot[p]) pOrgs += s.cls[ci].original();
            for (u32 ci : s.ot[n]) nOrgs += s.cls[ci].original();
        } else pOrgs = u32(s.ot[p].size()), nOrgs = u32(s.ot[n].size());
        variable_elimination(s, tid, x, pOrgs, nOrgs);
    }
    // phase 2: exclusive scans with the current CNF sizes as initial values (elimination.cu:82-92)
    std::vector<u32> cnt(s.ve_rpos);
    std::vector<u64> words(s.ve_rref);
    u32 accp = u32(s.cls.size());
    u64 accr = s.data_size;
    for (u32 t = 0; t < E; t++) {
        const u32 c = s.ve_rpos[t]; const u64 w = s.ve_rref[t];
        s.ve_rpos[t] = accp, s.ve_rref[t] = accr;
        accp += c, accr += w;
    }
    // phase 3 (ve_k_2)
    int lastEliminatedID = -1;
    std::vector<u32> survivors;
    for (u32 tid = 0; tid < E; tid++) {
        const u32 x = s.elected[tid];
        const u32 xinfo = s.ve_type[tid], elimType = xinfo & 3;
        if (elimType) {
            const u32 p = V2L(x), n = NEG(p);
            const u32 nAddedCls = (xinfo & 0xFFFC) >> 2, nAddedLits = (xinfo & 0xFFFF0000u) >> 16;
            const u32 addedPos = s.ve_rpos[tid];
            const u64 addedRef = s.ve_rref[tid];
            OL& poss = s.ot[p]; OL& negs = s.ot[n];
            const bool safe = (u64(addedPos) + nAddedCls <= s.refs_cap) &&
                              (addedRef + nAddedLits + u64(NBUCKETS) * nAddedCls <= s.data_cap);
            if (safe) {
                if (lastEliminatedID + 1 != int(tid) && s.cls.size() != addedPos) {
                    // an earlier variable failed MEMORY_SAFE: the reference leaves a hole here
                    fprintf(stderr, "oracle: unsupported hole in CNF after a failed MEMORY_SAFE\n"); abort();
                }
                saveResolved(s, p, n, poss.size() > negs.size(), poss, negs);
                emit_resolvents(s, x, elimType, nAddedCls, addedPos, addedRef, poss, negs);
                s.eliminated[x] |= MELTING_MASK | ADDING_MASK;
                lastEliminatedID = int(tid);
            } else if (elimType != RES_MASK) freezeClauses(s, poss, negs);
        }
        if (!s.eliminated[x]) survivors.push_back(x);
    }
    // postVE: resizeCNF_k cnf.cu:55-79
    if (lastEliminatedID >= 0) {
        const u32 info = s.ve_type[lastEliminatedID];
        const u32 cl = (info & 0xFFFC) >> 2, li = (info & 0xFFFF0000u) >> 16;
        s.data_size = s.ve_rref[lastEliminatedID] + li + u64(NBUCKETS) * cl;
        if (s.cls.size() != size_t(s.ve_rpos[lastEliminatedID]) + cl) { fprintf(stderr, "oracle: CNF size mismatch after BVE\n"); abort(); }
    }
    s.elected.swap(survivors);
}

// ------------------------------------------------------------------ SUB (subsume.cuh)
// subsume.cuh:50-75
bool sub(const S& s, const Clause& subsuming, const Clause& subsumed) {
    const u32* d1 = s.L(subsuming); const u32* e1 = d1 + subsuming.sz;
    const u32* d2 = s.L(subsumed); const u32* e2 = d2 + subsumed.sz;
    int n = 0;
    while (d1 != e1 && d2 != e2) {
        if (*d1 < *d2) d1++;
        else if (*d2 < *d1) d2++;
        else { n++; d1++, d2++; }
    }
    return n == subsuming.sz;
}
// subsume.cuh:132-176
bool selfsub_merge(const S& s, u32 x, u32 fx, const Clause& subsuming, const Clause& subsumed) {
    const u32* d1 = s.L(subsuming); const u32* e1 = d1 + subsuming.sz;
    const u32* d2 = s.L(subsumed); const u32* e2 = d2 + subsumed.sz;
    int n = 0; bool self = false;
    while (d1 != e1 && d2 != e2) {
        const u32 lit1 = *d1, lit2 = *d2;
        if (lit1 == fx) d1++;
        else if (lit2 == x) { self = true; d2++; }
        else if (lit1 < lit2) d1++;
        else if (lit2 < lit1) d2++;
        else { n++; d1++, d2++; }
    }
    if (n + 1 == subsuming.sz) {
        if (self) return true;
        while (d2 != e2) { if (*d2 == x) return true; d2++; }
    }
    return false;
}
inline bool selfsub_sig(u32 A, u32 B) {  // elimination.cuh:67-71
    const u32 B_tmp = B | ((B & 0xAAAAAAAAu) >> 1) | ((B & 0x55555555u) << 1);
    return !(A & ~B_tmp);
}
// subsume.cuh:226-238
void bumpShrunken(Clause& c) {
    const int old_lbd = int(c.lbd);
    if (old_lbd <= LBD_TIER1) return;
    const int new_lbd = std::min(c.sz - 1, old_lbd);
    if (new_lbd >= old_lbd) return;
    c.lbd = u32(new_lbd);
    c.usage = USAGET3;
}
// subsume.cuh:269-291
void strengthen(S& s, Clause& c, u32 self) {
    u32* l = s.L(c);
    int n = 0;
    for (int k = 0; k < c.sz; k++) if (l[k] != self) l[n++] = l[k];
    c.sz--;
    if (c.sz > 1) { calcSig(s, c); if (c.learnt()) bumpShrunken(c); }
}
// subsume.cuh:344-370
void selfsubsume(S& s, u32 x, u32 fx, const OL& list, Clause& cand, u32& nUnits) {
    const int candsz = cand.sz;
    const u32 candsig = cand.sig;
    for (u32 j : list) {
        const Clause& subsuming = s.cls[j];
        const int subsize = subsuming.sz;
        if (subsize > candsz) break;
        if (subsuming.deleted() || subsuming.molten) continue;
        if (subsize > 1 && selfsub_sig(subsuming.sig, candsig) && selfsub_merge(s, x, fx, subsuming, cand)) {
            strengthen(s, cand, x);
            cand.molten = 1;
            if (cand.sz == 1) nUnits++;
            break;
        }
    }
}
// subsume.cuh:305-323
void subsume(S& s, const OL& list, size_t end, Clause& cand) {
    const int candsz = cand.sz;
    for (size_t j = 0; j < end; j++) {
        Clause& subsuming = s.cls[list[j]];
        if (subsuming.deleted()) continue;
        if (cand.molten && subsuming.sz > candsz) continue;
        if (subsuming.sz > 1 && SUBSIG(subsuming.sig, cand.sig) && sub(s, subsuming, cand)) {
            if (subsuming.learnt() && cand.original()) subsuming.st = ORIGINAL;
            cand.st = DELETED;
            break;
        }
    }
}
// subsume.cuh:293-303
void updateOL(S& s, OL& ol) {
    size_t j = 0;
    for (size_t i = 0; i < ol.size(); i++) {
        Clause& c = s.cls[ol[i]];
        if (c.molten) c.molten = 0;
        else if (!c.deleted()) ol[j++] = ol[i];
    }
    ol.resize(j);
}
// saveProof (proofutils.cuh:182-200): the list filter of updateOL plus the proof lines - a strengthened
// (molten) clause is ADDED in its new form, a subsumed one DELETED; a clause that is both gets both lines
void saveProofSub(S& s, OL& ol) {
    size_t j = 0;
    for (size_t i = 0; i < ol.size(); i++) {
        Clause& c = s.cls[ol[i]];
        const bool deleted = c.deleted();
        if (c.molten) { saveProofClause(s, c, PROOF_ADDED); c.molten = 0; }
        else if (!deleted) ol[j++] = ol[i];
        if (deleted) saveProofClause(s, c, PROOF_DELETED);
    }
    ol.resize(j);
}
// subsume.cuh:402-484
void SUB(S& s) {
    for (u32 tid = 0; tid < s.numElected; tid++) {
        const u32 x = s.elected[tid], p = V2L(x), n = NEG(p);
        OL& poss = s.ot[p]; OL& negs = s.ot[n];
        if (poss.size() > s.o.sub_max_occurs || negs.size() > s.o.sub_max_occurs) continue;
        u32 nPosUnits = 0, nNegUnits = 0;
        for (size_t i = 0; i < poss.size(); i++) {
            Clause& pos = s.cls[poss[i]];
            if (pos.sz > SUB_MAX_CL_SIZE) break;
            if (pos.deleted()) continue;
            selfsubsume(s, p, n, negs, pos, nPosUnits);
            subsume(s, poss, i, pos);
        }
        for (size_t i = 0; i < negs.size(); i++) {
            Clause& neg = s.cls[negs[i]];
            if (neg.sz > SUB_MAX_CL_SIZE) break;
            if (neg.deleted()) continue;
            selfsubsume(s, n, p, poss, neg, nNegUnits);
            subsume(s, negs, i, neg);
        }
        if (nPosUnits || nNegUnits) {
            if (nPosUnits) appendUnits(s, poss);
            if (nNegUnits) appendUnits(s, negs);
        }
        if (s.proof_en) { saveProofSub(s, poss); saveProofSub(s, negs); }   // subsume.cuh:465-475
        else { updateOL(s, poss); updateOL(s, negs); }
    }
}

// ------------------------------------------------------------------ BCE (blocked.cuh:26-97)
void BCE(S& s) {
    for (u32 tid = 0; tid < s.elected.size(); tid++) {
        const u32 x = s.elected[tid], p = V2L(x), n = NEG(p);
        const OL& poss = s.ot[p]; const OL& negs = s.ot[n];
        if (poss.size() > s.o.bce_max_occurs || negs.size() > s.o.bce_max_occurs) continue;
        for (u32 i : negs) {
            Clause& ci = s.cls[i];
            if (ci.deleted() || ci.learnt()) continue;
            bool allTautology = true;
            for (u32 j : poss) {
                const Clause& cj = s.cls[j];
                if (cj.deleted() || cj.learnt()) continue;
                if (!isTautology(s, x, ci, cj)) { allTautology = false; break; }
            }
            if (allTautology) {
                saveClause(s, ci, n);
                if (s.proof_en) saveProofClause(s, ci, PROOF_DELETED);   // blocked.cuh:67-72
                ci.st = DELETED;
            }
        }
    }
}

// ------------------------------------------------------------------ ERE (redundancy.cuh:99-174)
void forward_equ(S& s, const u32* m_c, int m_len, u32 type) {
    u32 best = m_c[0], m_sig = MAPHASH(best);
    int minsize = int(s.ot[best].size());
    for (int k = 1; k < m_len; k++) {
        const u32 lit = m_c[k];
        const int lsize = int(s.ot[lit].size());
        if (lsize < minsize) minsize = lsize, best = lit;
        m_sig |= MAPHASH(lit);
    }
    const OL& minList = s.ot[best];
    // lane t checks entries t, t+32, ... and deletes its first match
    for (int lane = 0; lane < 32; lane++) {
        for (int i = lane; i < minsize; i += 32) {
            Clause& c = s.cls[minList[i]];
            if (m_len == c.sz && (c.learnt() || c.st == type) && SUBSIG(m_sig, c.sig) && !c.deleted() &&
                std::equal(m_c, m_c + m_len, s.L(c))) {
                c.st = DELETED;
                if (s.proof_en) saveProofClause(s, c, PROOF_DELETED);   // redundancy.cuh:122-129
                break;
            }
        }
    }
}
void ERE(S& s) {
    const int clause_max = s.o.ere_clause_max;
    std::vector<u32> m_c;
    for (u32 gid = 0; gid < s.numElected; gid++) {
        const u32 v = s.elected[gid], p = V2L(v), n = NEG(p);
        const OL& poss = s.ot[p]; const OL& negs = s.ot[n];
        const size_t ds = poss.size(), fs = negs.size();
        if (!(ds && fs && ds <= s.o.ere_max_occurs && fs <= s.o.ere_max_occurs &&
              s.cls[poss[0]].sz <= clause_max && s.cls[negs[0]].sz <= clause_max)) continue;
        for (u32 i : poss) {
            const Clause& pos = s.cls[i];
            if (pos.deleted()) continue;
            for (u32 j : negs) {
                const Clause& neg = s.cls[j];
                if (neg.deleted() || (pos.sz + neg.sz - 2) > clause_max) continue;
                m_c.resize(size_t(pos.sz) + neg.sz);
                const int m_len = merge_out(s, v, s.L(pos), pos.sz, s.L(neg), neg.sz, m_c.data());
                if (m_len > 1) {
                    const u32 type = (pos.learnt() || neg.learnt()) ? LEARNT : ORIGINAL;
                    forward_equ(s, m_c.data(), m_len, type);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ round loop
void snapshot(S& s) {
    Snapshot sn;
    sn.offs.push_back(0);
    for (const Clause& c : s.cls) {
        if (c.deleted()) continue;
        sn.bits.push_back(c.st | (c.molten << 2) | (c.added << 3) | (c.usage << 4) | (c.lbd << 6));
        sn.sig.push_back(c.sig);
        sn.lits.insert(sn.lits.end(), s.L(c), s.L(c) + c.sz);
        sn.offs.push_back(sn.lits.size());
    }
    s.snaps.push_back(std::move(sn));
}

// solver.hpp:748-753
bool stop(const S& s, i64 cr, i64 lr) {
    return (s.phase == s.o.phases) || (s.simpstate == CNFALLOC_FAIL) || (!cr && !lr) ||
           (s.phase > 2 && lr <= s.o.phase_lits_min);
}

// simplify.cu:136-241
void simplifying(S& s) {
    // awaken :77-134
    const u64 C0 = s.cls.size();
    u64 L0 = 0;
    for (const Clause& c : s.cls) L0 += c.sz;
    u64 numCls = C0, numLits = L0;
    if (s.o.phases) {
        numCls += s.o.ve_en ? s.orgClauses : 0;
        numLits += s.o.ve_en ? u64(double(s.orgLiterals) * s.o.lits_mul) : 0;
    }
    s.refs_cap = numCls;
    s.data_cap = numCls * NBUCKETS + numLits;
    s.numClauses = C0, s.numLiterals = L0;
    prepCNF(s);
    if (s.proof_en) {   // cuPROOF::count (proof.cu:101-121): bytes of every literal of the formula, x 1.5 (simplify.cu:128-132)
        u64 bytes = 0;
        for (const Clause& c : s.cls) for (int k = 0; k < c.sz; k++) bytes += proofLitBytes(s, s.L(c)[k]);
        s.proofCap = u32(double(u32(bytes)) * 1.5);
    }
    s.phase = s.multiplier = 0;
    i64 cdiff = INT64_MAX, ldiff = INT64_MAX;
    i64 clsbefore = i64(s.numClauses), litsbefore = i64(s.numLiterals);
    while (s.numClauses && s.numLiterals && !s.simpstate) {
        histSimp(s);                       // reallocOT
        reallocCNF(s);
        createOT(s);
        if (!prop(s)) return;              // UNSAT
        if (!s.numClauses) break;
        if (!LCVE(s)) break;
        sortOT(s);
        if (stop(s, cdiff, ldiff)) { if (s.o.ere_en && s.numElected) { ERE(s); flushProof(s); } break; }   // elimination.cu:305-306
        const u64 clsBeforeVE = s.cls.size();
        const u32 electedNow = s.numElected;
        if (s.o.sub_en || s.o.ve_plus_en) SUB(s);
        if (s.o.ve_en) VE(s);
        if (s.o.bce_en && !s.elected.empty()) BCE(s);
        flushProof(s);                     // cacheProof :174 ... writeProof :184
        u64 nc, nl;
        countAll(s, nc, nl);
        const u32 remained = u32(s.elected.size());
        s.numElected = remained;
        s.numClauses = nc, s.numLiterals = nl;
        cdiff = clsbefore - i64(nc), clsbefore = i64(nc);
        ldiff = litsbefore - i64(nl), litsbefore = i64(nl);
        s.nUnits = u32(s.units.size());
        s.phase++, s.multiplier++;
        s.multiplier += (s.phase == s.o.phases);
        s.rstats.push_back({electedNow, u64(electedNow - remained), u64(s.cls.size() - clsBeforeVE), nc, nl});
        if (s.keep_snaps) snapshot(s);
    }
    // write back :187-230
    if (s.unassigned <= 0 || !s.numClauses) { s.cnfstate = SAT; return; }
    if (s.o.final_gc && s.simpstate != CNFALLOC_FAIL && !s.compacted) {
        // reallocCNF(true) uses the counters of the last countAll (stale after ERE) only for
        // capacities; the clause list it produces is the live list either way
        u64 nc, nl; countAll(s, nc, nl);
        s.numClauses = nc, s.numLiterals = nl;
        compactCNF(s);
    }
    u64 nc, nl; countAll(s, nc, nl);
    s.numClauses = nc, s.numLiterals = nl;
}

} // namespace

// ==================================================================== C ABI
extern "C" {

void oracle_default_opts(oracle_opts* o) {
    memset(o, 0, sizeof *o);
    o->phases = 5; o->ve_en = 1; o->ve_plus_en = 1; o->sub_en = 1; o->bce_en = 0; o->ere_en = 1; o->all_en = 0;
    o->mu_pos = 32; o->mu_neg = 32; o->lcve_min_vars = 2; o->lcve_max_occurs = 3000; o->lcve_clause_max = 30000;
    o->phase_lits_min = 500; o->shrink_rate = 2; o->lits_mul = 1.0;
    o->ve_fun_en = 1; o->ve_lbound_en = 0; o->ve_clause_max = 100; o->xor_max_arity = 10;
    o->ere_clause_max = 250; o->ere_max_occurs = 3000; o->sub_max_occurs = 3000; o->bce_max_occurs = 3000;
    o->sh_max_bve_out1 = 250; o->sigma_calls = 1; o->final_gc = 1; o->aggr_cnf_sort = 0; o->lcve_fast = 0;
}

void oracle_normalize_opts(oracle_opts* o) {  // options.cpp:291-296
    o->ve_en = o->ve_en || o->ve_plus_en;
    if (o->all_en) o->ve_en = 1, o->ve_plus_en = 1, o->bce_en = 1, o->ere_en = 1;
    if (!o->phases && (o->ve_en || o->sub_en || o->bce_en)) o->phases = 1;
    if (o->phases && !(o->ve_en || o->sub_en || o->bce_en)) o->phases = 0;
    if (o->phases > 1 && !o->ve_en) o->phases = 1;
    if (o->ere_clause_max > 250) o->ere_clause_max = 250;
}

int oracle_create(const oracle_opts* o, uint32_t max_var, uint64_t num_clauses, const uint32_t* lits,
                  const uint64_t* offs, const uint32_t* meta, const uint32_t* vorg, const uint8_t* vstate,
                  oracle_ctx** out) {
    oracle_ctx* s = new oracle_ctx();
    s->o = *o;
    s->V = max_var;
    s->cls.resize(num_clauses);
    s->pool.assign(lits, lits + offs[num_clauses]);
    u64 ref = 0;
    for (u64 i = 0; i < num_clauses; i++) {
        Clause& c = s->cls[i];
        const u32 m = meta ? meta[i] : 0;
        c.st = m & 1; c.molten = 0; c.added = 0;
        c.usage = c.st ? ((m >> 4) & 3) : 0;
        c.lbd = c.st ? (m >> 6) : 0;
        c.sig = 0; c.sz = int(offs[i + 1] - offs[i]); c.off = offs[i]; c.ref = ref;
        ref += NBUCKETS + c.sz;
        if (c.st) {} else { s->orgClauses++; s->orgLiterals += c.sz; }
    }
    s->data_size = ref;
    const u32 nd = 2 * (max_var + 1);
    s->ot.resize(nd);
    s->hist.assign(nd, 0);
    s->eligible.assign(max_var + 1, 0);
    s->scores.assign(max_var + 1, 0);
    s->eliminated.assign(max_var + 1, 0);
    s->frozen.assign(max_var + 1, 0);
    s->vstate.assign(max_var + 1, 0);
    s->vorg.resize(max_var + 1);
    for (u32 v = 0; v <= max_var; v++) s->vorg[v] = vorg ? vorg[v] : v;
    s->unassigned = max_var;
    if (vstate) for (u32 v = 1; v <= max_var; v++) { s->vstate[v] = vstate[v]; if (vstate[v]) s->unassigned--; }
    *out = s;
    return 0;
}

int oracle_run(oracle_ctx* s) {
    const bool alldisabled = !s->o.phases && !(s->o.all_en | s->o.ere_en);  // solver.hpp:722
    if (!alldisabled && !s->cls.empty()) simplifying(*s);
    return s->cnfstate;
}

int oracle_rounds(const oracle_ctx* s) { return int(s->rstats.size()); }
void oracle_round_stats(const oracle_ctx* s, uint64_t* out) {
    for (size_t r = 0; r < s->rstats.size(); r++) {
        const RoundStat& t = s->rstats[r];
        out[5 * r + 0] = t.elected, out[5 * r + 1] = t.eliminated, out[5 * r + 2] = t.added;
        out[5 * r + 3] = t.clauses, out[5 * r + 4] = t.literals;
    }
}

static bool live_result(const oracle_ctx* s) { return s->cnfstate == UNSOLVED; }

uint64_t oracle_num_clauses(const oracle_ctx* s) {
    if (!live_result(s)) return 0;
    u64 n = 0; for (const Clause& c : s->cls) n += !c.deleted(); return n;
}
uint64_t oracle_num_literals(const oracle_ctx* s) {
    if (!live_result(s)) return 0;
    u64 n = 0; for (const Clause& c : s->cls) if (!c.deleted()) n += c.sz; return n;
}
uint64_t oracle_num_resolved(const oracle_ctx* s) { return s->resolved.size(); }
uint64_t oracle_num_trail(const oracle_ctx* s) { return s->trail.size(); }

void oracle_copy_result(const oracle_ctx* s, uint32_t* bits, uint32_t* sig, uint64_t* offs, uint32_t* lits,
                        uint8_t* eliminated, uint32_t* resolved, uint32_t* trail) {
    u64 i = 0, l = 0;
    offs[0] = 0;
    if (live_result(s)) {
        std::vector<u32> order;
        for (u32 k = 0; k < u32(s->cls.size()); k++) if (!s->cls[k].deleted()) order.push_back(k);
        // cacheCNF, cnf.cu:232-233: thrust::stable_sort of the refs with OLIST_CMP (key.cuh:67-83):
        // size, first literal, last literal, signature, ref
        if (s->o.aggr_cnf_sort && s->simpstate != OTALLOC_FAIL && s->simpstate != CNFALLOC_FAIL)   // !reallocFailed()
            std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) {
                const Clause& x = s->cls[a]; const Clause& y = s->cls[b];
                if (x.sz != y.sz) return x.sz < y.sz;
                const u32 *lx = s->L(x), *ly = s->L(y);
                if (lx[0] != ly[0]) return lx[0] < ly[0];
                if (lx[x.sz - 1] != ly[y.sz - 1]) return lx[x.sz - 1] < ly[y.sz - 1];
                if (x.sig != y.sig) return x.sig < y.sig;
                return a < b;
            });
        for (u32 k : order) {
            const Clause& c = s->cls[k];
            bits[i] = c.st | (c.molten << 2) | (c.added << 3) | (c.usage << 4) | (c.lbd << 6);
            sig[i] = c.sig;
            memcpy(lits + l, s->L(c), size_t(c.sz) * 4);
            l += c.sz;
            offs[++i] = l;
        }
    }
    memcpy(eliminated, s->eliminated.data(), s->eliminated.size());
    if (!s->resolved.empty()) memcpy(resolved, s->resolved.data(), s->resolved.size() * 4);
    if (!s->trail.empty()) memcpy(trail, s->trail.data(), s->trail.size() * 4);
}

void oracle_keep_snapshots(oracle_ctx* s, int keep) { s->keep_snaps = keep != 0; }
void oracle_enable_proof(oracle_ctx* s, int on) { s->proof_en = on != 0; }
int oracle_proof_chunks(const oracle_ctx* s) { return int(s->proofChunks.size()); }
uint64_t oracle_proof_chunk_size(const oracle_ctx* s, int i) { return s->proofChunks[size_t(i)].size(); }
void oracle_copy_proof_chunk(const oracle_ctx* s, int i, uint8_t* out) {
    const std::vector<uint8_t>& c = s->proofChunks[size_t(i)];
    if (!c.empty()) memcpy(out, c.data(), c.size());
}
uint32_t oracle_proof_capacity(const oracle_ctx* s) { return s->proofCap; }
// incremental solving: variables under assumption are never candidates (Solver::LCVE, lcve.cu:316-323, lcve_k :88)
void oracle_set_assumed(oracle_ctx* s, const uint8_t* assumed) {
    if (assumed) s->assumed.assign(assumed, assumed + s->V + 1); else s->assumed.clear();
}
uint64_t oracle_snapshot_clauses(const oracle_ctx* s, int r) { return s->snaps[r].bits.size(); }
uint64_t oracle_snapshot_literals(const oracle_ctx* s, int r) { return s->snaps[r].lits.size(); }
void oracle_copy_snapshot(const oracle_ctx* s, int r, uint32_t* bits, uint32_t* sig, uint64_t* offs, uint32_t* lits) {
    const Snapshot& sn = s->snaps[r];
    if (!sn.bits.empty()) { memcpy(bits, sn.bits.data(), sn.bits.size() * 4); memcpy(sig, sn.sig.data(), sn.sig.size() * 4); }
    memcpy(offs, sn.offs.data(), sn.offs.size() * 8);
    if (!sn.lits.empty()) memcpy(lits, sn.lits.data(), sn.lits.size() * 4);
}

int oracle_num_elections(const oracle_ctx* s) { return int(s->electedLog.size()); }
uint64_t oracle_election_size(const oracle_ctx* s, int i) { return s->electedLog[i].size(); }
void oracle_copy_election(const oracle_ctx* s, int i, uint32_t* out) {
    if (!s->electedLog[i].empty()) memcpy(out, s->electedLog[i].data(), s->electedLog[i].size() * 4);
}

int oracle_write_dump(const oracle_ctx* s, const char* path) {
    const u64 nc = oracle_num_clauses(s), nl = oracle_num_literals(s);
    std::vector<u32> bits(nc), sig(nc), lits(nl), resolved(s->resolved.size()), trail(s->trail.size());
    std::vector<u64> offs(nc + 1);
    std::vector<uint8_t> elim(s->V + 1);
    oracle_copy_result(s, bits.data(), sig.data(), offs.data(), lits.data(), elim.data(), resolved.data(), trail.data());
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    const u32 nelim = s->V + 1;
    const u32 hdr[12] = {0x31444753u, s->V, u32(s->cnfstate), u32(nc), u32(3 * nc + nl), nelim, u32(resolved.size()),
                         u32(trail.size()), u32(s->numClauses), u32(s->numLiterals), u32(s->simpstate), 0};
    fwrite(hdr, 4, 12, f);
    for (u64 i = 0; i < nc; i++) {
        const u32 h[3] = {bits[i], sig[i], u32(offs[i + 1] - offs[i])};
        fwrite(h, 4, 3, f);
        fwrite(lits.data() + offs[i], 4, offs[i + 1] - offs[i], f);
    }
    elim.resize(size_t((nelim + 3) / 4) * 4, 0);
    fwrite(elim.data(), 1, elim.size(), f);
    if (!resolved.empty()) fwrite(resolved.data(), 4, resolved.size(), f);
    if (!trail.empty()) fwrite(trail.data(), 4, trail.size(), f);
    return fclose(f);
}

void oracle_destroy(oracle_ctx* s) { delete s; }

void oracle_prep(uint64_t num_clauses, uint32_t* lits, const uint64_t* offs, uint32_t* sig) {
    for (u64 i = 0; i < num_clauses; i++) {
        u32* b = lits + offs[i]; u32* e = lits + offs[i + 1];
        std::sort(b, e);
        u32 sg = 0;
        if (e - b > 1) for (u32* k = b; k != e; k++) sg |= MAPHASH(*k);
        sig[i] = sg;
    }
}

void oracle_histogram(uint64_t num_lits, const uint32_t* lits, uint32_t nbins, uint32_t* hist) {
    memset(hist, 0, size_t(nbins) * 4);
    for (u64 i = 0; i < num_lits; i++) hist[lits[i]]++;
}

// model.cpp:101-162
uint64_t oracle_extend_model(uint8_t* value, uint32_t max_var, const uint32_t* resolved, uint64_t n) {
    (void)max_var;
    u64 updated = 0;
    if (!n) return 0;
    const u32* x = resolved + n - 1;
    while (x > resolved) {
        bool unsat = true;
        u32 k;
        for (k = *x--; k > 1; k--, x--) {
            if (value[ABS(*x)] == !SIGN(*x)) { unsat = false; break; }
        }
        if (unsat) { value[ABS(*x)] = !SIGN(*x); updated++; }
        x -= k;
    }
    return updated;
}

uint64_t oracle_check_model(const uint8_t* value, uint64_t num_clauses, const uint32_t* lits, const uint64_t* offs) {
    u64 bad = 0;
    for (u64 i = 0; i < num_clauses; i++) {
        bool sat = false;
        for (u64 k = offs[i]; k < offs[i + 1] && !sat; k++) sat = value[ABS(lits[k])] == !SIGN(lits[k]);
        bad += !sat;
    }
    return bad;
}

// ---- helpers of tests/sgd.py (fingerprints of 100 M-word witness stacks in seconds instead of minutes of Python loops)
// Records of the witness stack are `[lits..., size]` (model.cuh:29-53): the sizes sit at the END, so the boundaries are found
// walking back from the top.  ends[k] = one past record k, ascending; returns the number of records, ~0 if the stack is corrupt.
uint64_t oracle_record_ends(const uint32_t* r, uint64_t n, uint64_t* ends) {
    uint64_t cnt = 0;
    for (uint64_t p = n; p > 0;) { const uint64_t sz = r[p - 1]; if (!sz || sz + 1 > p) return ~0ull; p -= sz + 1; cnt++; }
    if (ends) { uint64_t k = cnt; for (uint64_t p = n; p > 0;) { ends[--k] = p; p -= (uint64_t)r[p - 1] + 1; } }
    return cnt;
}
// FNV-1a over the words of every segment [starts[k], ends[k]) (sgd.hash_words)
void oracle_hash_segments(const uint32_t* r, const uint64_t* starts, const uint64_t* ends, uint64_t m, uint64_t* out) {
    for (uint64_t k = 0; k < m; k++) {
        uint64_t h = 0xCBF29CE484222325ull;
        for (uint64_t p = starts[k]; p < ends[k]; p++) { h ^= r[p]; h *= 0x100000001B3ull; }
        out[k] = h;
    }
}

} // extern "C"

#ifdef ORACLE_MAIN
// CLI: sigma_oracle <in.cnf> <out.sgd> [--phases=K] [-no-ere] [-no-vefunction] [-bce] [-all] [-no-sub] [-no-veextend]
#include <string>
static bool read_dimacs(const char* path, u32& V, std::vector<u32>& lits, std::vector<u64>& offs) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> buf(size_t(n) + 1);
    if (fread(buf.data(), 1, size_t(n), f) != size_t(n)) { fclose(f); return false; }
    fclose(f);
    buf[size_t(n)] = 0;
    char* p = buf.data();
    offs.assign(1, 0);
    V = 0;
    while (*p) {
        while (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t') p++;
        if (!*p) break;
        if (*p == 'c') { while (*p && *p != '\n') p++; continue; }
        if (*p == 'p') { p += 5; V = u32(strtoul(p, &p, 10)); strtoul(p, &p, 10); continue; }
        const long v = strtol(p, &p, 10);
        if (v == 0) offs.push_back(lits.size());
        else lits.push_back(v < 0 ? (u32(-v) << 1) | 1 : u32(v) << 1);
    }
    return true;
}
int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <in.cnf> <out.sgd> [flags]\n", argv[0]); return 2; }
    oracle_opts o; oracle_default_opts(&o);
    for (int i = 3; i < argc; i++) {
        const std::string a = argv[i];
        if (a.rfind("--phases=", 0) == 0) o.phases = atoi(a.c_str() + 9);
        else if (a == "-no-ere") o.ere_en = 0;
        else if (a == "-no-vefunction") o.ve_fun_en = 0;
        else if (a == "-bce") o.bce_en = 1;
        else if (a == "-all") o.all_en = 1;
        else if (a == "-no-sub") o.sub_en = 0;
        else if (a == "-no-veextend") o.ve_plus_en = 0;
        else if (a == "-no-ve") o.ve_en = 0;
        else if (a == "-no-lcvefast" || a == "-quiet") {}
        else if (a == "-lcvefast") o.lcve_fast = 1;
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    oracle_normalize_opts(&o);
    u32 V; std::vector<u32> lits; std::vector<u64> offs;
    if (!read_dimacs(argv[1], V, lits, offs)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    oracle_ctx* c;
    oracle_create(&o, V, offs.size() - 1, lits.data(), offs.data(), nullptr, nullptr, nullptr, &c);
    const int st = oracle_run(c);
    printf("c oracle: state %d, rounds %d, clauses %llu, literals %llu\n", st, oracle_rounds(c),
           (unsigned long long)oracle_num_clauses(c), (unsigned long long)oracle_num_literals(c));
    const int rc = oracle_write_dump(c, argv[2]);
    oracle_destroy(c);
    return rc;
}
#endif
