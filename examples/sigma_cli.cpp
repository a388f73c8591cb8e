// examples/sigma_cli.cpp -- the C ABI (include/sigma.h) from a plain C++ program, no reference code involved:
// DIMACS in, one SIGmA call on the GPU, simplified DIMACS + witness stack (+ binary DRAT with -proof) out.
//
//   g++ -O2 -std=c++17 -Iinclude examples/sigma_cli.cpp -Lparafrost_b200 -lsigma_b200 -Wl,-rpath,$PWD/parafrost_b200 -o build/sigma_cli
//   build/sigma_cli in.cnf [--out simplified.cnf] [--witness w.txt] [--proof p.drat] [--phases=K] [-no-ere] [-bce] [-all] [-no-vefunction]
//
// Mirrors what `parafrost -no-solve` does around Solver::simplify (src/gpu/solver.cpp:207): parse, simplify, report.
// There is no CPU path: without a usable CUDA device sigma_create fails and the program says so.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sigma.h"

static bool readDimacs(const char* path, uint32_t& maxVar, std::vector<uint32_t>& lits, std::vector<uint64_t>& offs) {
    FILE* f = fopen(path, "r");
    if (!f) return false;
    maxVar = 0;
    offs.assign(1, 0);
    std::vector<uint32_t> cur;
    char tok[64];
    int ch;
    while ((ch = fgetc(f)) != EOF) {
        if (ch == 'c' || ch == 'p') { while ((ch = fgetc(f)) != EOF && ch != '\n') {} continue; }
        if (ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r') continue;
        int n = 0;
        tok[n++] = (char)ch;
        while ((ch = fgetc(f)) != EOF && ch != ' ' && ch != '\n' && ch != '\t' && ch != '\r' && n < 62) tok[n++] = (char)ch;
        tok[n] = 0;
        const long v = strtol(tok, nullptr, 10);
        if (v == 0) {
            if (!cur.empty()) { lits.insert(lits.end(), cur.begin(), cur.end()); offs.push_back(lits.size()); cur.clear(); }
        } else {
            const uint32_t var = (uint32_t)(v < 0 ? -v : v);
            if (var > maxVar) maxVar = var;
            cur.push_back(2 * var + (v < 0 ? 1u : 0u));      // lit = 2 var + sign (constants.hpp:72-80)
        }
    }
    fclose(f);
    return maxVar > 0 && offs.size() > 1;
}

static void proofToFile(void* user, const uint8_t* bytes, uint64_t n) { fwrite(bytes, 1, (size_t)n, (FILE*)user); }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.cnf [--out f] [--witness f] [--proof f] [reference simplifier flags]\n", argv[0]); return 2; }
    sigma_opts o;
    sigma_default_opts(&o);
    const char *out = nullptr, *wit = nullptr, *proof = nullptr;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--out" && i + 1 < argc) out = argv[++i];
        else if (a == "--witness" && i + 1 < argc) wit = argv[++i];
        else if (a == "--proof" && i + 1 < argc) { proof = argv[++i]; o.proof_en = 1; }
        else if (a.rfind("--phases=", 0) == 0) o.phases = atoi(a.c_str() + 9);
        else if (a == "-no-ere") o.ere_en = 0;
        else if (a == "-bce") o.bce_en = 1;
        else if (a == "-all") o.all_en = 1;
        else if (a == "-no-vefunction") o.ve_fun_en = 0;
        else if (a == "-no-sub") o.sub_en = 0;
        else if (a == "-profilegpu") o.profile = 1;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    sigma_normalize_opts(&o);
    uint32_t V = 0;
    std::vector<uint32_t> lits;
    std::vector<uint64_t> offs;
    if (!readDimacs(argv[1], V, lits, offs)) { fprintf(stderr, "cannot read a formula from %s\n", argv[1]); return 2; }
    printf("c %s: %u variables, %zu clauses, %zu literals\n", argv[1], V, offs.size() - 1, lits.size());
    sigma_ctx* c = nullptr;
    int rc = sigma_create(0, &o, &c);
    if (rc) { fprintf(stderr, "sigma_create failed (%d): no usable CUDA device - this engine has no CPU path\n", rc); return 1; }
    FILE* pf = nullptr;
    if (proof) { pf = fopen(proof, "wb"); if (!pf) { perror(proof); return 2; } sigma_set_proof_sink(c, proofToFile, pf); }
    sigma_report rep;
    rc = sigma_load(c, V, offs.size() - 1, lits.data(), offs.data(), nullptr, nullptr, nullptr, nullptr);
    if (!rc) rc = sigma_run(c, &rep);
    if (rc) { fprintf(stderr, "simplification failed (%d): %s\n", rc, sigma_last_error(c)); sigma_destroy(c); return 1; }
    if (pf) fclose(pf);
    uint64_t nC, nL, nR, nT;
    sigma_result_sizes(c, &nC, &nL, &nR, &nT);
    std::vector<uint32_t> bits(nC), sig(nC), olits(nL), resolved(nR), trail(nT);
    std::vector<uint64_t> ooffs(nC + 1);
    std::vector<uint8_t> elim(V + 1);
    sigma_store(c, bits.data(), sig.data(), ooffs.data(), olits.data(), elim.data(), resolved.data(), trail.data());
    printf("c %u rounds, %.3f ms on the device, %llu kernel launches\n", rep.rounds, rep.ms_device, (unsigned long long)rep.kernel_launches);
    printf("c eliminated %u variables, %llu -> %llu clauses, %llu -> %llu literals, %llu units, %llu witness words\n", rep.eliminated_vars,
           (unsigned long long)rep.clauses_in, (unsigned long long)nC, (unsigned long long)rep.literals_in, (unsigned long long)nL,
           (unsigned long long)nT, (unsigned long long)nR);
    printf("s %s\n", rep.cnfstate == SIGMA_UNSAT ? "UNSATISFIABLE" : rep.cnfstate == SIGMA_SAT ? "SATISFIABLE (by simplification)" : "UNKNOWN");
    if (out) {
        FILE* f = fopen(out, "w");
        if (!f) { perror(out); return 2; }
        fprintf(f, "p cnf %u %llu\n", V, (unsigned long long)(nC + nT));
        for (uint64_t i = 0; i < nT; i++) fprintf(f, "%s%u 0\n", (trail[i] & 1) ? "-" : "", trail[i] >> 1);
        for (uint64_t i = 0; i < nC; i++) {
            for (uint64_t k = ooffs[i]; k < ooffs[i + 1]; k++) fprintf(f, "%s%u ", (olits[k] & 1) ? "-" : "", olits[k] >> 1);
            fprintf(f, "0\n");
        }
        fclose(f);
    }
    if (wit) {   // the witness stack in the reference's format (model.cuh:29-53): groups of literals followed by their count, witness first
        FILE* f = fopen(wit, "w");
        if (!f) { perror(wit); return 2; }
        for (uint64_t i = 0; i < nR; i++) fprintf(f, "%u\n", resolved[i]);
        fclose(f);
    }
    sigma_destroy(c);
    return 0;
}
