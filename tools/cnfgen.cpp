// tools/cnfgen.cpp -- deterministic synthetic CNF generators for the BASELINE.json configs.
//
// There is no network and the reference ships no benchmark inputs, so every parity/bench
// input is generated in-box from a fixed seed.  PRNG = splitmix64,
// so the same (family, params, seed) gives the same CNF on every box.  No generated CNF
// contains unit clauses, duplicate literals or tautologies, so what the reference parser
// (src/gpu/dimacs.cpp) hands to Solver::simplify() is exactly the clause list written here.
//
// Literal encoding is the reference's (src/gpu/constants.hpp:72-80): lit = 2*var + sign.
//
// Built two ways from this one file:
//   g++ -O2 -shared -fPIC            -> libcnfgen.so   (C ABI below, used through ctypes)
//   g++ -O2 -DCNFGEN_MAIN            -> cnfgen         (CLI: writes DIMACS and/or CSR files)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    // unbiased enough for generators: 64-bit multiply-high
    uint32_t below(uint32_t n) { return uint32_t((__uint128_t(next()) * n) >> 64); }
    bool coin() { return next() >> 63; }
};

struct Cnf {
    uint32_t nvars = 0;
    std::vector<uint32_t> lits;
    std::vector<uint64_t> offs{0};
    uint32_t newvar() { return ++nvars; }
    void add(std::initializer_list<uint32_t> c) {
        for (uint32_t l : c) lits.push_back(l);
        offs.push_back(lits.size());
    }
    void add(const std::vector<uint32_t>& c) {
        lits.insert(lits.end(), c.begin(), c.end());
        offs.push_back(lits.size());
    }
    size_t nclauses() const { return offs.size() - 1; }
};

inline uint32_t P(uint32_t v) { return v << 1; }
inline uint32_t N(uint32_t v) { return (v << 1) | 1; }
inline uint32_t NOT(uint32_t l) { return l ^ 1; }

// g = a & b   (3 clauses)
void gate_and(Cnf& f, uint32_t g, uint32_t a, uint32_t b) {
    f.add({NOT(g), a});
    f.add({NOT(g), b});
    f.add({g, NOT(a), NOT(b)});
}
// g = a | b
void gate_or(Cnf& f, uint32_t g, uint32_t a, uint32_t b) {
    f.add({g, NOT(a)});
    f.add({g, NOT(b)});
    f.add({NOT(g), a, b});
}
// g = a ^ b   (4 clauses)
void gate_xor(Cnf& f, uint32_t g, uint32_t a, uint32_t b) {
    f.add({NOT(g), a, b});
    f.add({NOT(g), NOT(a), NOT(b)});
    f.add({g, NOT(a), b});
    f.add({g, a, NOT(b)});
}
// g = a ^ b ^ c (8 clauses)
void gate_xor3(Cnf& f, uint32_t g, uint32_t a, uint32_t b, uint32_t c) {
    for (int m = 0; m < 8; m++) {
        // forbid assignments where g != a^b^c: clause falsified exactly by that assignment
        const int va = m & 1, vb = (m >> 1) & 1, vc = (m >> 2) & 1;
        const int vg = va ^ vb ^ vc;
        // assignment (a=va,b=vb,c=vc,g=!vg) is forbidden
        f.add({vg ? g : NOT(g), va ? NOT(a) : a, vb ? NOT(b) : b, vc ? NOT(c) : c});
    }
}
// g = maj(a,b,c) (6 clauses)
void gate_maj(Cnf& f, uint32_t g, uint32_t a, uint32_t b, uint32_t c) {
    f.add({g, NOT(a), NOT(b)});
    f.add({g, NOT(a), NOT(c)});
    f.add({g, NOT(b), NOT(c)});
    f.add({NOT(g), a, b});
    f.add({NOT(g), a, c});
    f.add({NOT(g), b, c});
}
// g <-> l  (2 binary clauses)
void gate_equ(Cnf& f, uint32_t g, uint32_t l) {
    f.add({NOT(g), l});
    f.add({g, NOT(l)});
}
// force literal l true without a unit clause: (l | t)(l | !t)
void force(Cnf& f, uint32_t l, uint32_t t) {
    f.add({l, P(t)});
    f.add({l, N(t)});
}

// ---------------------------------------------------------------- families
void gen_ksat(Cnf& f, uint32_t n, uint64_t m, int k, uint64_t seed) {
    Rng r(seed);
    f.nvars = n;
    f.lits.reserve(m * k);
    f.offs.reserve(m + 1);
    std::vector<uint32_t> c(k);
    for (uint64_t i = 0; i < m; i++) {
        for (int j = 0; j < k;) {
            const uint32_t v = 1 + r.below(n);
            bool dup = false;
            for (int q = 0; q < j; q++) dup |= (c[q] >> 1) == v;
            if (dup) continue;
            c[j++] = (v << 1) | uint32_t(r.coin());
        }
        f.add(c);
    }
}

// Tseitin miter of two copies of a random AND/XOR circuit (BMC-like).
void gen_miter(Cnf& f, uint32_t ninputs, uint32_t ngates, uint32_t xor_permille,
               uint32_t rewrite_permille, uint32_t nouts, uint64_t seed) {
    Rng r(seed);
    struct G { uint8_t op; uint32_t a, b; }; // operands = literal over node ids of copy
    std::vector<G> gates(ngates);
    const uint32_t total = ninputs + ngates;
    for (uint32_t g = 0; g < ngates; g++) {
        const uint32_t avail = ninputs + g;
        uint32_t a = r.below(avail), b = r.below(avail);
        while (b == a) b = r.below(avail);
        gates[g].op = r.below(1000) < xor_permille;
        gates[g].a = (a << 1) | uint32_t(r.coin());
        gates[g].b = (b << 1) | uint32_t(r.coin());
    }
    // inputs are shared: vars 1..ninputs
    f.nvars = ninputs;
    std::vector<uint32_t> nodeA(total), nodeB(total);
    for (uint32_t i = 0; i < ninputs; i++) nodeA[i] = nodeB[i] = i + 1;
    auto lit_of = [](const std::vector<uint32_t>& node, uint32_t l) { return (node[l >> 1] << 1) | (l & 1); };
    for (uint32_t g = 0; g < ngates; g++) {
        const uint32_t out = f.newvar();
        nodeA[ninputs + g] = out;
        const uint32_t a = lit_of(nodeA, gates[g].a), b = lit_of(nodeA, gates[g].b);
        if (gates[g].op) gate_xor(f, P(out), a, b); else gate_and(f, P(out), a, b);
    }
    for (uint32_t g = 0; g < ngates; g++) {
        const uint32_t out = f.newvar();
        nodeB[ninputs + g] = out;
        const uint32_t a = lit_of(nodeB, gates[g].a), b = lit_of(nodeB, gates[g].b);
        if (r.below(1000) < rewrite_permille) {
            // equivalent form through an auxiliary node: out = !t, t = xnor / nand-as-or
            const uint32_t t = f.newvar();
            if (gates[g].op) gate_xor(f, P(t), NOT(a), b);       // t = xnor(a,b)
            else gate_or(f, P(t), NOT(a), NOT(b));               // t = !a | !b
            gate_equ(f, P(out), N(t));
        } else {
            if (gates[g].op) gate_xor(f, P(out), a, b); else gate_and(f, P(out), a, b);
        }
    }
    if (nouts > ngates) nouts = ngates;
    std::vector<uint32_t> big;
    for (uint32_t j = 0; j < nouts; j++) {
        const uint32_t d = f.newvar();
        gate_xor(f, P(d), P(nodeA[total - 1 - j]), P(nodeB[total - 1 - j]));
        big.push_back(P(d));
    }
    if (big.size() >= 2) f.add(big);
}

// n x n array multiplier; product bits tied to a seeded product of two odd n-bit numbers.
void gen_mult(Cnf& f, uint32_t n, uint64_t seed) {
    Rng r(seed);
    std::vector<uint32_t> a(n), b(n);
    for (auto& v : a) v = f.newvar();
    for (auto& v : b) v = f.newvar();
    // seeded factors (odd, top bit set) and their schoolbook product
    std::vector<uint8_t> fa(n), fb(n), prod(2 * n, 0);
    for (uint32_t i = 0; i < n; i++) fa[i] = r.coin(), fb[i] = r.coin();
    fa[0] = fb[0] = 1; fa[n - 1] = fb[n - 1] = 1;
    {
        std::vector<uint32_t> acc(2 * n + 1, 0);
        for (uint32_t i = 0; i < n; i++) if (fa[i]) for (uint32_t j = 0; j < n; j++) acc[i + j] += fb[j];
        uint32_t carry = 0;
        for (uint32_t k = 0; k < 2 * n; k++) { const uint32_t s = acc[k] + carry; prod[k] = s & 1; carry = s >> 1; }
    }
    // partial products
    std::vector<std::vector<uint32_t>> pp(n, std::vector<uint32_t>(n));
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = 0; j < n; j++) {
            pp[i][j] = f.newvar();
            gate_and(f, P(pp[i][j]), P(a[i]), P(b[j]));
        }
    // row-by-row ripple array: sum[k] holds the running sum bit of weight k
    std::vector<uint32_t> p(2 * n, 0);
    std::vector<uint32_t> sum(n);                 // weights i..i+n-1 of the running row
    for (uint32_t j = 0; j < n; j++) sum[j] = pp[0][j];
    p[0] = sum[0];
    uint32_t top = 0;                             // carry-out of previous row (weight i+n-1), 0 = none
    for (uint32_t i = 1; i < n; i++) {
        // add pp[i][0..n-1] (weights i..i+n-1) to (sum[1..n-1], top) (weights i..i+n-1)
        std::vector<uint32_t> nsum(n);
        uint32_t carry = 0;
        for (uint32_t j = 0; j < n; j++) {
            const uint32_t x = pp[i][j];
            const uint32_t y = (j + 1 < n) ? sum[j + 1] : top;
            if (!y && !carry) { nsum[j] = x; continue; }
            const uint32_t s = f.newvar(), c = f.newvar();
            if (y && carry) { gate_xor3(f, P(s), P(x), P(y), P(carry)); gate_maj(f, P(c), P(x), P(y), P(carry)); }
            else { const uint32_t z = y ? y : carry; gate_xor(f, P(s), P(x), P(z)); gate_and(f, P(c), P(x), P(z)); }
            nsum[j] = s; carry = c;
        }
        sum.swap(nsum);
        top = carry;
        p[i] = sum[0];
    }
    for (uint32_t j = 1; j < n; j++) p[n - 1 + j] = sum[j];
    p[2 * n - 1] = top;
    const uint32_t t = f.newvar();
    for (uint32_t k = 0; k < 2 * n; k++) if (p[k]) force(f, prod[k] ? P(p[k]) : N(p[k]), t);
}

// x1 ^ ... ^ xn = b, Tseitin chained (4 clauses per link)
void gen_parity(Cnf& f, uint32_t n, uint64_t seed) {
    Rng r(seed);
    std::vector<uint32_t> x(n);
    for (auto& v : x) v = f.newvar();
    uint32_t acc = x[0];
    for (uint32_t i = 1; i < n; i++) {
        const uint32_t tnew = f.newvar();
        gate_xor(f, P(tnew), P(acc), P(x[i]));
        acc = tnew;
    }
    const uint32_t t = f.newvar();
    force(f, r.coin() ? P(acc) : N(acc), t);
}

bool generate(Cnf& f, const std::string& fam, const std::vector<uint64_t>& a, uint64_t seed) {
    auto A = [&](size_t i, uint64_t d) { return i < a.size() ? a[i] : d; };
    if (fam == "ksat") gen_ksat(f, uint32_t(A(0, 1000)), A(1, 4260), int(A(2, 3)), seed);
    else if (fam == "miter") gen_miter(f, uint32_t(A(0, 100)), uint32_t(A(1, 1000)), uint32_t(A(2, 900)), uint32_t(A(3, 100)), uint32_t(A(4, 32)), seed);
    else if (fam == "mult") gen_mult(f, uint32_t(A(0, 8)), seed);
    else if (fam == "parity") gen_parity(f, uint32_t(A(0, 100)), seed);
    else if (fam == "multpar") { gen_mult(f, uint32_t(A(0, 8)), seed); gen_parity(f, uint32_t(A(1, 100)), seed ^ 0x5bd1e995u); }
    else return false;
    return true;
}

bool write_dimacs(const Cnf& f, const char* path) {
    FILE* out = fopen(path, "wb");
    if (!out) return false;
    std::vector<char> buf;
    buf.reserve(1 << 22);
    char tmp[64];
    int n = snprintf(tmp, sizeof tmp, "p cnf %u %zu\n", f.nvars, f.nclauses());
    buf.insert(buf.end(), tmp, tmp + n);
    for (size_t c = 0; c < f.nclauses(); c++) {
        for (uint64_t k = f.offs[c]; k < f.offs[c + 1]; k++) {
            const uint32_t l = f.lits[k];
            char* q = tmp + sizeof tmp;
            *--q = ' ';
            uint32_t v = l >> 1;
            do { *--q = char('0' + v % 10); v /= 10; } while (v);
            if (l & 1) *--q = '-';
            buf.insert(buf.end(), q, tmp + sizeof tmp);
        }
        buf.push_back('0'); buf.push_back('\n');
        if (buf.size() > (1u << 22) - 4096) { fwrite(buf.data(), 1, buf.size(), out); buf.clear(); }
    }
    fwrite(buf.data(), 1, buf.size(), out);
    return fclose(out) == 0;
}

} // namespace

// ---------------------------------------------------------------- C ABI (ctypes)
extern "C" {

struct cnfgen_handle { Cnf f; };

// family: "ksat" (n, m, k) | "miter" (inputs, gates, xor_permille, rewrite_permille, outs)
//       | "mult" (bits) | "parity" (n) | "multpar" (bits, n)
int cnfgen_create(const char* family, const uint64_t* args, int nargs, uint64_t seed, cnfgen_handle** out) {
    cnfgen_handle* h = new cnfgen_handle();
    std::vector<uint64_t> a(args, args + nargs);
    if (!generate(h->f, family, a, seed)) { delete h; return 1; }
    *out = h;
    return 0;
}
uint32_t cnfgen_nvars(const cnfgen_handle* h) { return h->f.nvars; }
uint64_t cnfgen_nclauses(const cnfgen_handle* h) { return h->f.nclauses(); }
uint64_t cnfgen_nlits(const cnfgen_handle* h) { return h->f.lits.size(); }
void cnfgen_copy(const cnfgen_handle* h, uint32_t* lits, uint64_t* offs) {
    memcpy(lits, h->f.lits.data(), h->f.lits.size() * sizeof(uint32_t));
    memcpy(offs, h->f.offs.data(), h->f.offs.size() * sizeof(uint64_t));
}
int cnfgen_write_dimacs(const cnfgen_handle* h, const char* path) { return write_dimacs(h->f, path) ? 0 : 1; }
void cnfgen_destroy(cnfgen_handle* h) { delete h; }

}

#ifdef CNFGEN_MAIN
int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <family> <seed> <out.cnf> [args...]\n", argv[0]);
        return 2;
    }
    std::vector<uint64_t> a;
    for (int i = 4; i < argc; i++) a.push_back(strtoull(argv[i], nullptr, 10));
    Cnf f;
    if (!generate(f, argv[1], a, strtoull(argv[2], nullptr, 10))) { fprintf(stderr, "unknown family %s\n", argv[1]); return 2; }
    if (!write_dimacs(f, argv[3])) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
    printf("c cnfgen %s seed %s: vars %u clauses %zu literals %zu\n", argv[1], argv[2], f.nvars, f.nclauses(), f.lits.size());
    return 0;
}
#endif
