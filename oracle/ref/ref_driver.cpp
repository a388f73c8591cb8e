/* oracle/ref/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * A replacement `main` for the *unmodified* reference GPU solver
 * (/root/reference/src/gpu, compiled where it lies by oracle/ref/Makefile).
 * It drives the reference's own public entry points
 *   Solver::Solver(path)            (src/gpu/solver.cpp:42)
 *   Solver::simplify(skip=true)     (src/gpu/simplify.cu:57)
 * and then serialises what the simplifier left behind, so that the CPU
 * restatement in oracle/ and the CUDA engine can be compared with it bit
 * for bit.  No reference source is copied here: this file only *calls* the
 * reference through its headers.
 *
 * Output ("SGD1" dump, little endian uint32 words unless noted):
 *   magic 'SGD1', maxVar, cnfstate, nClauses, nDataWords, nElim(=maxVar+1),
 *   nResolved, nTrail, numClauses(inf), numLiterals(inf), simpstate, pad
 *   data words of every live clause in ref order, each clause being the
 *     reference's SCLAUSE record: {bits(st:2,f:1,a:1,u:2,lbd:26), sig, size, lits[size]}
 *   eliminated bytes padded to a word multiple
 *   resolved words (model.resolved, src/gpu/model.hpp)
 *   trail words (root-level units, includes those enqueued by Solver::prop)
 *
 * Usage: ref_driver <cnf> <dump-out> [reference CLI flags...]
 */
#include "control.hpp"
#include "banner.hpp"
#include "solver.hpp"
#include "options.cuh"
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cstdlib>
#include <chrono>

using namespace ParaFROST;

bool quiet_en = false;
int  verbose = -1;

namespace {

struct RefDriver : public Solver {
	explicit RefDriver(const std::string& path) : Solver(path) {}

	int run(const char* out_path)
	{
		initLimits();
		const bool hostMode = getenv("REF_DRIVER_HOST") != NULL;
		const auto t0 = std::chrono::steady_clock::now();
		if (canPreSimplify()) simplify(!hostMode);
		const cudaError_t syncErr = cudaDeviceSynchronize();
		const cudaError_t lastErr = cudaGetLastError();
		const auto t1 = std::chrono::steady_clock::now();
		const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		if (syncErr != cudaSuccess || lastErr != cudaSuccess)
			printf("c ref_driver: CUDA ERROR after simplify: sync=%d (%s) last=%d (%s)\n",
				int(syncErr), cudaGetErrorString(syncErr), int(lastErr), cudaGetErrorString(lastErr));

		std::vector<uint32> words;
		uint32 nClauses = 0;
		if (hostMode && IS_UNSOLVED(cnfstate)) {
			// simplify(false) rebuilt the host clause database (newBeginning -> writeBackCNF ->
			// newClause, src/gpu/cnf.cu:186-198): dump orgs then learnts, literals re-sorted
			// because watch handling may have permuted them. sig/added are not kept by the host.
			BCNF* sets[2] = { &orgs, &learnts };
			for (int k = 0; k < 2; k++) {
				BCNF& set = *sets[k];
				for (uint32 i = 0; i < set.size(); i++) {
					CLAUSE& c = cm[set[i]];
					if (c.deleted()) continue;
					std::vector<uint32> l(c.data(), c.data() + c.size());
					std::sort(l.begin(), l.end());
					uint32 bits = c.learnt() ? 1u : 0u;
					if (c.learnt()) bits |= (uint32(c.usage()) << 4) | (uint32(c.lbd()) << 6);
					words.push_back(bits); words.push_back(0); words.push_back(uint32(c.size()));
					words.insert(words.end(), l.begin(), l.end());
					nClauses++;
				}
			}
		}
		else if (hcnf && IS_UNSOLVED(cnfstate)) {
			for (uint32 i = 0; i < hcnf->size(); i++) {
				SCLAUSE& c = hcnf->clause(i);
				if (c.deleted()) continue;
				const uint32* raw = (const uint32*)&c;
				const uint32 n = SCLAUSEBUCKETS + uint32(c.size());
				words.insert(words.end(), raw, raw + n);
				nClauses++;
			}
		}
		// witnesses: simplify(true) does not cache them unless the CNF got emptied
		if (vars && IS_UNSOLVED(cnfstate) && !hostMode) { cacheResolved(streams[2]); cudaDeviceSynchronize(); }
		const uint32 nElim = inf.maxVar + 1;
		std::vector<Byte> elim(nElim, 0);
		if (vars) {
			if (vars->isEliminatedCached && vars->cachedEliminated)
				for (uint32 v = 0; v < nElim; v++) elim[v] = vars->cachedEliminated[v];
			else
				cudaMemcpy(elim.data(), vars->eliminated, nElim, cudaMemcpyDeviceToHost);
		}
		const uint32 elimWords = (nElim + 3) / 4;
		elim.resize(size_t(elimWords) * 4, 0);

		FILE* f = fopen(out_path, "wb");
		if (!f) { fprintf(stderr, "ref_driver: cannot open %s\n", out_path); return 2; }
		const uint32 hdr[12] = {
			0x31444753u /* 'SGD1' */, inf.maxVar, uint32(cnfstate), nClauses, uint32(words.size()),
			nElim, model.resolved.size(), trail.size(), inf.numClauses, inf.numLiterals,
			uint32(simpstate), uint32(lastErr != cudaSuccess ? lastErr : syncErr) };
		fwrite(hdr, sizeof(uint32), 12, f);
		if (!words.empty()) fwrite(words.data(), sizeof(uint32), words.size(), f);
		fwrite(elim.data(), 1, elim.size(), f);
		if (model.resolved.size()) fwrite(model.resolved.data(), sizeof(uint32), model.resolved.size(), f);
		if (trail.size()) fwrite(trail.data(), sizeof(uint32), trail.size(), f);
		fclose(f);
		printf("c ref_driver: simplify wall %.3f ms, state %d, clauses %u, data words %zu, resolved %u, trail %u\n",
			ms, int(cnfstate), nClauses, words.size(), model.resolved.size(), trail.size());
		if (gopts.profile_gpu) {
			printf("c ref_driver: stage ms vo %.3f sig %.3f io %.3f gc %.3f cot %.3f sot %.3f rot %.3f ve %.3f sub %.3f bce %.3f ere %.3f\n",
				stats.sigma.time.vo, stats.sigma.time.sig, stats.sigma.time.io, stats.sigma.time.gc,
				stats.sigma.time.cot, stats.sigma.time.sot, stats.sigma.time.rot, stats.sigma.time.ve,
				stats.sigma.time.sub, stats.sigma.time.bce, stats.sigma.time.ere);
		}
		fflush(stdout);
		return 0;
	}
};

}

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s <cnf> <dump-out> [flags]\n", argv[0]); return 2; }
	BOOL_OPT opt_quiet_en("quiet", "enable quiet mode, same as verbose=0", false);
	INT_OPT opt_verbose("verbose", "set the verbosity", 1, INT32R(0, 4));
	// the reference parser treats argv[1] as the formula and the rest as flags
	std::vector<char*> args;
	args.push_back(argv[0]);
	args.push_back(argv[1]);
	for (int i = 3; i < argc; i++) args.push_back(argv[i]);
	int nargs = int(args.size());
	try {
		parseArguments(nargs, args.data());
		quiet_en = opt_quiet_en, verbose = opt_verbose;
		if (quiet_en) verbose = 0;
		else if (!verbose) quiet_en = true;
		signal_handler(handler_terminate);
		RefDriver* drv = new RefDriver(std::string(argv[1]));
		init_solver(drv);
		const int rc = drv->run(argv[2]);
		fflush(NULL); // stdout and, with -proof, the reference's proof file (PROOF::write is buffered stdio)
		_exit(rc); // skip the reference's teardown; the dump is on disk
	}
	catch (std::bad_alloc&) { fprintf(stderr, "ref_driver: bad_alloc\n"); return 3; }
	catch (MEMOUTEXCEPTION&) { fprintf(stderr, "ref_driver: memout\n"); return 3; }
}
