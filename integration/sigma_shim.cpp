/* integration/sigma_shim.cpp -- the reference-side binding of libsigma_b200.
 *
 * ParaFROST has no plugin API: its simplifier is a set of `Solver` members (src/gpu/solver.hpp:674-789)
 * whose definitions live in the reference's 16 CUDA translation units.  The reference's 45 host objects
 * import exactly seven symbols from them (nm on the objects built by oracle/ref/Makefile):
 *     Solver::simplify(bool const&)   Solver::optSimp()   Solver::freeSimp()   Solver::newBeginning()
 *     cuMM::cuMM()                    CACHER::destroy()   GOPTION::GOPTION()
 * This translation unit defines those seven on top of the C ABI in include/sigma.h, so that
 *     reference host objects + this file + libsigma_b200.so     (integration/Makefile)
 * link into a `parafrost` whose CDCL, parser, CLI, proof file and model extension are the reference's own
 * and whose inprocessing runs on the B200 engine.  It includes the reference's headers and is therefore
 * compiled next to the reference (never into libsigma_b200.so, whose sources include none of them).
 *
 * What each body replaces is cited beside it.  Nothing here is copied from the reference's .cu files: the
 * host-side steps of simplifying() (simplify.cu:136-241) are re-expressed through the public/protected
 * members of `Solver` that the host objects already define.
 */
#include "solver.hpp"
#include "options.cuh"
#include "memory.cuh"

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "sigma.h"

using namespace ParaFROST;

// ---- the simplifier's command-line options (options.cu:33-60): the objects register themselves with the
// reference's option parser (input.hpp), so `parafrost -h` and the flags keep working unchanged
BOOL_OPT opt_sync_always_en("syncalways", "accepted for compatibility (the engine orders its work on one stream)", false);
BOOL_OPT opt_profile_gpu_en("profilegpu", "per-stage CUDA-event times of the simplifier", false);
INT_OPT opt_ve_min_threads("veminthreads", "accepted for compatibility (no effect: group-per-variable kernels)", 4, INT32R(2, 1024));
INT_OPT opt_sub_min_threads("subminthreads", "accepted for compatibility (no effect)", 4, INT32R(2, 1024));
INT_OPT opt_ere_min_threads("ereminthreads", "accepted for compatibility (no effect)", 4, INT32R(2, 1024));
DOUBLE_OPT opt_ve_min_blocks("veminblocks", "accepted for compatibility (no effect)", 0.5, FP64R(0, 1));
DOUBLE_OPT opt_sub_min_blocks("subminblocks", "accepted for compatibility (no effect)", 0.5, FP64R(0, 1));
DOUBLE_OPT opt_ere_min_blocks("ereminblocks", "accepted for compatibility (no effect)", 0.5, FP64R(0, 1));
BOOL_OPT opt_ve_fun_en("vefunction", "function-table reasoning in BVE", true);
BOOL_OPT opt_ve_lbound_en("velitsbound", "skip eliminations that add more literals than they remove", false);
INT_OPT opt_bce_max_occurs("bcemaxoccurs", "longest occurrence list scanned in BCE", 3e3, INT32R(100, INT32_MAX));
INT_OPT opt_sub_max_occurs("submaxoccurs", "longest occurrence list scanned in SUB", 3e3, INT32R(100, INT32_MAX));
INT_OPT opt_sub_clause_max("subclausemax", "accepted for compatibility", 100, INT32R(0, INT32_MAX));
INT_OPT opt_ere_extend("ereextend", "accepted for compatibility", 1, INT32R(0, 3));
INT_OPT opt_ere_max_occurs("eremaxoccurs", "longest occurrence list scanned in ERE", 3e3, INT32R(100, INT32_MAX));
INT_OPT opt_ere_clause_max("ereclausemax", "largest resolvent checked in ERE", 250, INT32R(2, INT32_MAX));
INT_OPT opt_ve_clause_max("resolventmax", "largest resolvent BVE may add (0: no limit)", 100, INT32R(0, INT32_MAX));
INT_OPT opt_xor_max_arity("xormaxarity", "largest XOR gate looked for", 10, INT32R(2, 20));

namespace ParaFROST {
	__constant__ KOptions kOpts[1];     // declared extern in options.cuh; the engine passes its options by value
	void initDevOpts() {}
}

// ---- the four inert symbols: the engine owns all device memory (one arena per context)
GOPTION::GOPTION() { RESETSTRUCT(this); }                       // options.cu:62
void GOPTION::init(const bool& proof_en)                         // options.cu:66-89
{
	sync_always = opt_sync_always_en;
	profile_gpu = opt_profile_gpu_en;
	ve_min_threads = opt_ve_min_threads, sub_min_threads = opt_sub_min_threads, ere_min_threads = opt_ere_min_threads;
	ve_min_blocks = opt_ve_min_blocks, sub_min_blocks = opt_sub_min_blocks, ere_min_blocks = opt_ere_min_blocks;
	hostKOpts.proof_en = proof_en;
	hostKOpts.ve_fun_en = opt_ve_fun_en;
	hostKOpts.ve_lbound_en = opt_ve_lbound_en;
	hostKOpts.ve_clause_max = opt_ve_clause_max;
	hostKOpts.xor_max_arity = opt_xor_max_arity;
	hostKOpts.ere_clause_max = MIN(int(opt_ere_clause_max), 250);
	hostKOpts.ere_max_occurs = opt_ere_max_occurs;
	hostKOpts.sub_max_occurs = opt_sub_max_occurs;
	hostKOpts.sub_clause_max = opt_sub_clause_max;
	hostKOpts.bce_max_occurs = opt_bce_max_occurs;
}
cuMM::cuMM() :                                                   // memory.cu:85-97 (never used: nothing is allocated through it)
	pinned_cnf(nullptr), d_refs_mem(nullptr), d_scatter(nullptr), d_segs(nullptr), d_occurs(nullptr), d_hist(nullptr),
	d_cnf_mem(nullptr), d_stencil(nullptr), d_vstate(nullptr), nscatters(0), _compacttime(0.0f), cap(0), dcap(0), penalty(0) {}
void CACHER::destroy() {}                                        // cache.cu

// ---- state of the binding (the reference is one Solver per process)
namespace {
	sigma_ctx*             g_ctx = nullptr;
	size_t                 g_nData = 0, g_nRefs = 0;   // the simplified CNF as the reference's SCLAUSE record stream, in the pinned buffers below
	std::vector<uint32_t>  g_resolved;     // witness stack produced by this call (cacheResolved, transfer.cu:83-96)

	void fillOpts(sigma_opts& o, const OPTION& opts, const GOPTION& g, const int calls)
	{
		sigma_default_opts(&o);
		o.phases = opts.phases;
		o.ve_en = opts.ve_en, o.ve_plus_en = opts.ve_plus_en, o.sub_en = opts.sub_en;
		o.bce_en = opts.bce_en, o.ere_en = opts.ere_en, o.all_en = opts.all_en;
		o.mu_pos = opts.mu_pos, o.mu_neg = opts.mu_neg;
		o.lcve_min_vars = opts.lcve_min_vars, o.lcve_max_occurs = opts.lcve_max_occurs, o.lcve_clause_max = opts.lcve_clause_max;
		o.phase_lits_min = opts.phase_lits_min, o.shrink_rate = opts.shrink_rate, o.lits_mul = opts.lits_mul;
		o.ve_fun_en = g.hostKOpts.ve_fun_en, o.ve_lbound_en = g.hostKOpts.ve_lbound_en;
		o.ve_clause_max = g.hostKOpts.ve_clause_max, o.xor_max_arity = g.hostKOpts.xor_max_arity;
		o.ere_clause_max = g.hostKOpts.ere_clause_max, o.ere_max_occurs = g.hostKOpts.ere_max_occurs;
		o.sub_max_occurs = g.hostKOpts.sub_max_occurs, o.bce_max_occurs = g.hostKOpts.bce_max_occurs;
		o.sigma_calls = calls;
		o.final_gc = 1;
		o.profile = g.profile_gpu;
		o.aggr_cnf_sort = opts.aggr_cnf_sort;
		o.proof_en = g.hostKOpts.proof_en;
		o.log_reductions = verbose >= 2;              // LOGREDALL / LOGREDCL (logging.hpp:152-158)
		o.lcve_fast = opts.lcve_fast;                 // -lcvefast, the CLI's default (options.cpp:32): filtered-candidate MIS
		sigma_normalize_opts(&o);
	}

	// pinned host buffers of the two edges (cuMM::createMirror keeps the reference's mirror pinned as well, memory.cu)
	struct Pinned {
		void* p = nullptr; size_t cap = 0;
		void* get(size_t bytes) {
			if (bytes > cap) { sigma_pinned_free(p); cap = bytes + bytes / 8 + 4096; p = sigma_pinned_alloc(cap); if (!p) cap = 0; }
			return p;
		}
		void release() { sigma_pinned_free(p); p = nullptr; cap = 0; }
	};
	Pinned g_pinData, g_pinRefs;

	// runs fn(begin, end, worker) over [0, n) on the host cores (extractCNF / markEliminated are per-item independent)
	template <class F> void parallelFor(size_t n, F fn)
	{
		unsigned hw = std::thread::hardware_concurrency();
		size_t T = hw ? hw : 4;
		if (T > 32) T = 32;
		if (n < 65536 || T == 1) { fn(size_t(0), n, 0u); return; }
		std::vector<std::thread> th;
		const size_t per = (n + T - 1) / T;
		for (size_t t = 0; t < T; t++) {
			const size_t b = t * per, e = std::min(n, b + per);
			if (b >= e) break;
			th.emplace_back([=] { fn(b, e, unsigned(t)); });
		}
		for (auto& x : th) x.join();
	}

	// units of one device prop() in the reference's host order (elimbcp.cu:185-200): the first `seeds` came from SUB/BVE
	// (enqueueDevUnit: frozen, no proof line), the rest were derived (enqueueUnit: learnt, proof line).  A literal that is
	// on the trail already (two variables can emit the same unit) is skipped instead of being pushed twice.
	struct UnitReplay { Solver* solver; uint64_t seen; };
	UnitReplay g_replay = { nullptr, 0 };

}

// a member of Solver so that it can reach enqueueDevUnit / enqueueUnit: replays the units of the last device prop()
void Solver::cacheUnits(const cudaStream_t&)
{
	uint64_t total = 0; uint32_t from = 0, count = 0, seeds = 0;
	if (sigma_trail_info(g_ctx, &total, &from, &count, &seeds) != SIGMA_OK) return;
	if (uint64_t(from) + count <= g_replay.seen || !count) return;       // this prop() was replayed already
	std::vector<uint32_t> units(count);
	if (sigma_copy_trail(g_ctx, from, count, units.data()) != SIGMA_OK) return;
	for (uint32_t i = 0; i < count; i++) {
		const uint32 unit = units[i];
		if (!active(unit)) continue;
		if (i < seeds) enqueueDevUnit(unit);
		else enqueueUnit(unit);
	}
	sp->propagated = trail.size();
	stats.units.forced += count;
	g_replay.seen = uint64_t(from) + count;
}

namespace {
	// cuPROOF::writeProof (proof.cu:160-199): the round's binary DRAT bytes go to the reference's proof file - after the
	// units of the round's prop(), whose lines the reference writes at the top of the round
	void proofSink(void* user, const uint8_t* bytes, uint64_t n)
	{
		Solver* solver = static_cast<Solver*>(user);
		solver->cacheUnits(0);
		for (uint64_t i = 0; i < n; i++) solver->proof.write(Byte(bytes[i]));
	}
}

// Solver::optSimp (simplify.cu:243-252): device options + the engine's context on device 0
void Solver::optSimp()
{
	gopts.init(opts.proof_en && !opts.proof_nonbinary_en);
	sigma_opts o;
	fillOpts(o, opts, gopts, 1);
	if (g_ctx == nullptr && sigma_create(0, &o, &g_ctx) != SIGMA_OK) {
		LOGERRORN("cannot create the simplifier context on device 0");
		killSolver();
	}
	if (o.proof_en) sigma_set_proof_sink(g_ctx, proofSink, this);
}

// Solver::freeSimp (simplify.cu:254-266)
void Solver::freeSimp()
{
	if (g_ctx != nullptr) sigma_destroy(g_ctx), g_ctx = nullptr;
	g_pinData.release(), g_pinRefs.release();
	g_nData = g_nRefs = 0, g_resolved.clear();
	vars = NULL, cnf = NULL, ot = NULL, hcnf = NULL;
}

// Solver::newBeginning (transfer.cu:25-40): rebuild the host clause database from the simplified CNF.
// Reached directly or through map(true) (vmap.cpp:117), which sets `mapped` so that newClause renumbers.
void Solver::newBeginning()
{
	assert(wt.empty());
	assert(orgs.empty());
	assert(learnts.empty());
	cm.init(g_nData);
	// cacheResolved (transfer.cu:83-96)
	if (!g_resolved.empty()) {
		const uint32 off = model.resolved.size();
		model.resolved.resize(off + uint32(g_resolved.size()));
		uint32* start = model.resolved + off;
		for (size_t i = 0; i < g_resolved.size(); i++) start[i] = g_resolved[i];
		g_resolved.clear();
	}
	// writeBackCNF (cnf.cu:186-198): the records are the reference's own SCLAUSE layout (sclause.cuh:37-42)
	stats.literals.original = stats.literals.learnt = 0;
	uint32_t* data = static_cast<uint32_t*>(g_pinData.p);
	const uint64_t* refs = static_cast<const uint64_t*>(g_pinRefs.p);
	for (size_t i = 0; i < g_nRefs; i++) {      // serial: newClause allocates in the reference's clause arena `cm`
		SCLAUSE& s = *reinterpret_cast<SCLAUSE*>(data + refs[i]);
		if (s.deleted()) continue;
		newClause(s);
	}
	stats.clauses.original = orgs.size();
	stats.clauses.learnt = learnts.size();
	g_nData = g_nRefs = 0;
}

// Solver::simplify + Solver::simplifying (simplify.cu:57-75, 136-241)
void Solver::simplify(const bool& skip_transfer_to_host)
{
	if (alldisabled()) return;
	assert(conflict == NOREF);
	assert(IS_UNSOLVED(cnfstate));
	stats.sigma.calls++;
	do {
		SLEEPING(sleep.sigma, opts.sigma_sleep_en);
		rootify();
		shrinkTop(false);
		if (orgs.empty()) { recycleWT(); break; }
		timer.stop();
		stats.time.solve += timer.cpuTime();
		timer.start();
		// ---- awaken (simplify.cu:77-134).  extractCNF (cnf.cu:176-184) in two parallel passes: sizes, prefix sum, then every
		// worker writes its clauses as the reference's own SCLAUSE records {word 0, sig, size, literals} straight into pinned
		// memory - the mirror `hcnf` that reflectCNF ships (cnf.cu:166-174), taken by sigma_load_sclauses as it is.
		simpstate = AWAKEN_SUCC;
		const size_t nOrg = orgs.size(), nAll = nOrg + learnts.size();
		std::vector<uint64_t> pos(nAll + 1);                      // word offset of clause i's record, then compacted to refs
		parallelFor(nAll, [&](size_t b, size_t e, unsigned) {
			for (size_t i = b; i < e; i++) {
				CLAUSE& c = cm[i < nOrg ? orgs[uint32(i)] : learnts[uint32(i - nOrg)]];
				pos[i] = c.deleted() ? 0 : uint64_t(c.size()) + 3;
			}
		});
		uint64_t words = 0, nCls = 0;
		for (size_t i = 0; i < nAll; i++) { const uint64_t w = pos[i]; pos[i] = words; words += w; nCls += w != 0; }
		pos[nAll] = words;
		uint32_t* hdata = static_cast<uint32_t*>(g_pinData.get((words + 4) * sizeof(uint32_t)));
		uint64_t* hrefs = static_cast<uint64_t*>(g_pinRefs.get((nAll + 2) * sizeof(uint64_t)));
		if (!hdata || !hrefs) { simpstate = AWAKEN_FAIL; recycle(); break; }
		std::vector<uint64_t> rank(nAll + 1);                     // index of clause i among the live ones
		{ uint64_t r = 0; for (size_t i = 0; i < nAll; i++) { rank[i] = r; r += pos[i + 1] != pos[i]; } rank[nAll] = r; }
		std::atomic<uint64_t> nLits(0);
		parallelFor(nAll, [&](size_t b, size_t e, unsigned) {
			uint64_t lits = 0;
			for (size_t i = b; i < e; i++) {
				if (pos[i + 1] == pos[i]) continue;
				CLAUSE& c = cm[i < nOrg ? orgs[uint32(i)] : learnts[uint32(i - nOrg)]];
				uint32_t* rec = hdata + pos[i];
				rec[0] = c.learnt() ? (1u | (uint32_t(c.usage()) << 4) | (uint32_t(c.lbd()) << 6)) : 0u;
				rec[1] = 0;
				rec[2] = uint32_t(c.size());
				for (int j = 0; j < c.size(); j++) rec[3 + j] = c[j];
				hrefs[rank[i]] = pos[i];
				lits += uint64_t(c.size());
			}
			nLits += lits;
		});
		std::vector<uint8_t> vstate(inf.maxVar + 1, 0), assumedMask(inf.maxVar + 1, 0);
		forall_variables(v) {
			vstate[v] = uint8_t(sp->vstate[v].state);
			assumedMask[v] = iassumed(v) ? 1 : 0;                    // lcve.cu:316-323
		}
		sigma_opts o;
		fillOpts(o, opts, gopts, int(stats.sigma.calls));
		int rc = sigma_set_opts(g_ctx, &o);
		if (!rc) rc = sigma_load_sclauses(g_ctx, inf.maxVar, nCls, hdata, words, hrefs, vorg.data(), vstate.data(),
		                                  incremental ? assumedMask.data() : nullptr);
		if (rc) {                                                   // simplify.cu:152-155
			LOGWARNING("simplifier could not load the formula (%d: %s)", rc, sigma_last_error(g_ctx));
			simpstate = (rc == SIGMA_AWAKEN_FAIL) ? AWAKEN_FAIL : CNFALLOC_FAIL;
			recycle();
			break;
		}
		printStats(1, '-', CGREEN0);
		wt.clear(true), orgs.clear(true), learnts.clear(true);
		cm.destroy();
		// ---- reduction phases (simplify.cu:156-186) on the device, one loop iteration per call: the units of a round's
		// prop() reach the host (and the proof file) before that round's proof chunk and before the next round
		const int64 imelted = inf.maxMelted, iclauses = int64(nCls), iliterals = int64(nLits.load());
		g_replay.solver = this, g_replay.seen = 0;
		rc = sigma_begin(g_ctx);
		int done = 0;
		uint32_t redSeen = 0;
		while (!rc && !done) {
			rc = sigma_round(g_ctx, nullptr, &done);
			if (!rc) cacheUnits(0);                                 // elimbcp.cu:185-200 (no-op when the proof sink did it already)
			if (!rc && verbose >= 2) {                              // LOGREDALL / LOGREDCL tables of this round (count.cu:185-208)
				static const char* names[5] = { "BCP Reductions", "SUB Reductions", "BVE Reductions", "BCE Reductions", "ERE Reductions" };
				uint32_t n = 0;
				sigma_reduction_log(g_ctx, nullptr, &n);
				std::vector<sigma_stage_reduction> red(n ? n : 1);
				uint32_t cap = n;
				sigma_reduction_log(g_ctx, red.data(), &cap);
				for (uint32_t i = redSeen; i < n; i++) {
					const sigma_stage_reduction& e = red[i];
					LOG1("\t\t %s%s%s", CLBLUE, names[e.stage < 5 ? e.stage : 0], CNORMAL);
					LOG1("  %s%-10s  %-10s %-10s %-10s%s", CREPORT, " ", "Variables", "Clauses", "Literals", CNORMAL);
					LOG1("  %s%-10s: %s-%-8u  -%-8lld  -%-8lld%s", CREPORT, "Removed", CREPORTVAL, e.vars_removed,
						(long long)(e.clauses_before - e.clauses), (long long)(e.literals_before - e.literals), CNORMAL);
					LOG1("  %s%-10s: %s%-10s  %-9llu  %-9llu%s", CREPORT, "Survived", CREPORTVAL, " ",
						(unsigned long long)e.clauses, (unsigned long long)e.literals, CNORMAL);
				}
				redSeen = n;
			}
		}
		sigma_report rep;
		if (!rc) rc = sigma_finish(g_ctx, &rep);
		if (rc) {
			LOGERRORN("simplifier failed (%d: %s)", rc, sigma_last_error(g_ctx));
			killSolver();
		}
		uint64_t nC = 0, nL = 0, nR = 0, nT = 0;
		sigma_result_sizes(g_ctx, &nC, &nL, &nR, &nT);
		std::vector<uint8_t> eliminated(inf.maxVar + 1, 0);
		g_resolved.resize(nR);
		sigma_store_compact(g_ctx, nullptr, nullptr, nullptr, eliminated.data(), g_resolved.data(), nullptr);
		if (rep.cnfstate == SIGMA_UNSAT) { learnEmpty(); killSolver(); }   // elimbcp.cu:178, simplify.cu:168
		inf.numClauses = uint32(rep.clauses), inf.numLiterals = uint32(rep.literals);
		// ---- write back (simplify.cu:187-240)
		inf.maxMelted += rep.eliminated_vars;
		const bool success = (iclauses != int64(inf.numClauses));
		stats.sigma.all.clauses += iclauses - int64(inf.numClauses);
		stats.sigma.all.literals += iliterals - int64(inf.numLiterals);
		stats.sigma.all.variables += int64(inf.maxMelted) - imelted;
		last.shrink.removed = stats.shrunken;
		if (!inf.unassigned || !inf.numClauses || rep.cnfstate == SIGMA_SAT) {   // simplify.cu:198-209 (before markEliminated, as there)
			const uint32 off = model.resolved.size();
			model.resolved.resize(off + uint32(g_resolved.size()));
			for (size_t i = 0; i < g_resolved.size(); i++) model.resolved[off + uint32(i)] = g_resolved[i];
			g_resolved.clear();
			stats.clauses.original = stats.clauses.learnt = 0;
			stats.literals.original = stats.literals.learnt = 0;
			cnfstate = SAT;
			printStats(1, 's', CGREEN);
			break;
		}
		{                                                                // markEliminated (transfer.cu:42-60, solver.hpp:521-527) over the host cores
			std::atomic<uint32_t> melted(0);
			VSTATE* vs = sp->vstate;
			parallelFor(size_t(inf.maxVar), [&](size_t b, size_t e, unsigned) {
				uint32_t m = 0;
				for (size_t v = b + 1; v <= e; v++)
					if (eliminated[v] && !IS_FORCED(eliminated[v])) { vs[v].state = MELTED_M; m++; }
				melted += m;
			});
			inf.unassigned -= melted.load();
		}
		if (skip_transfer_to_host) {                                     // the simplified CNF stays resident in the context (sigma_device_view, sigma_continue)
			printStats(1, 's', CGREEN);
			break;
		}
		g_nData = size_t(3) * nC + nL, g_nRefs = nC;
		uint32_t* odata = static_cast<uint32_t*>(g_pinData.get((g_nData + 4) * sizeof(uint32_t)));
		uint64_t* orefs = static_cast<uint64_t*>(g_pinRefs.get((g_nRefs + 2) * sizeof(uint64_t)));
		if (!odata || !orefs) { LOGERRORN("no pinned memory for the write-back"); killSolver(); }
		sigma_store_sclauses(g_ctx, odata, orefs);                       // cacheCNF, cnf.cu:200-237
		if (canMap()) map(true);
		else newBeginning();
		rebuildWT(opts.sigma_priorbins);
		if (BCP()) {
			LOG2(1, " Propagation after simplify proved a contradiction");
			learnEmpty();
		}
		UPDATE_SLEEPER(this, sigma, success);
		printStats(1, 's', CGREEN);
		timer.stop(), stats.time.simp += timer.cpuTime();
	} while (0);
	INCREASE_LIMIT(this, sigma, stats.sigma.calls, nlognlogn, true);
	last.sigma.reduces = stats.reduces + 1;
	if (opts.phases > 2) opts.phases--;
	if (!opts.solve_en) killSolver();
	timer.start();
}
