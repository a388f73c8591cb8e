"""GPU parity tests: the CUDA engine (through the C ABI, include/sigma.h) against the CPU oracle
on the same seeded inputs, against the committed golden dumps of the unmodified reference, and
- at BASELINE.json's full sizes - through size-independent properties.

Bit-exact bar (integer work): same elected counts, same eliminated set, same resolvent counts,
same clause list *in the same order with the same flag and signature words* after every round,
same witness groups, same units; every model extended from the witness stack satisfies the
original CNF."""
import json
import os

import numpy as np
import pytest

import helpers
import sgd

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SUMMARY = json.load(open(os.path.join(HERE, "golden", "summary.json")))

SMALL = {
    "k3_r30": ("ksat", 12, [800, 2400, 3]),
    "k3_r42": ("ksat", 11, [600, 2520, 3]),
    "k5_r10": ("ksat", 13, [400, 4000, 5]),
    "k4_r7": ("ksat", 14, [500, 3500, 4]),
    "miter_x": ("miter", 21, [30, 600, 900, 100, 8]),
    "miter_a": ("miter", 22, [40, 700, 300, 200, 8]),
    "mult6": ("mult", 31, [6]),
    "mult10": ("mult", 32, [10]),
    "parity": ("parity", 41, [300]),
    "multpar": ("multpar", 51, [5, 120]),
}
MEDIUM = {
    "cfg1_k3_100k": ("ksat", 1, [100000, 426000, 3]),      # BASELINE config 1, full size
    "miter_50k": ("miter", 3, [2000, 50000, 900, 100, 32]),
    "mult48": ("mult", 4, [48]),
    "k5_20k": ("ksat", 2, [20000, 200000, 5]),
}
VARIANTS = {
    "p5": ["--phases=5", "-no-ere"],
    "def": [],
    "nofun": ["-no-vefunction"],
    "all": ["-all"],
    "bce": ["-bce"],
    "nosub_p2": ["-no-sub", "-no-veextend", "--phases=2", "-no-ere"],
}


def sigma():
    from parafrost_b200 import sigma as s
    return s


def to_dump(V, st, state=2):
    return sgd.Dump.from_arrays(V, state, st["bits"], st["sig"], st["offs"], st["lits"], st["eliminated"], st["resolved"], st["trail"])


def run_engine_rounds(V, lits, offs, flags, oracle_stats=None, oracle_snaps=None, elections=None, meta=None):
    """Drives sigma_begin/round/finish and checks every round against the oracle's."""
    s = sigma().Simplifier(0, flags=flags)
    try:
        s.load(V, lits, offs, meta=meta)
        s.begin()
        r = 0
        while True:
            rep, done = s.round()
            if rep["kind"] == 0:
                if oracle_stats is not None:
                    assert r < len(oracle_stats), f"engine ran an extra round {rep}"
                    got = (rep["elected"], rep["eliminated"], rep["resolvents"], rep["clauses"], rep["literals"])
                    assert got == tuple(int(x) for x in oracle_stats[r]), f"round {r}"
                if oracle_snaps is not None:
                    snap = to_dump(V, s.snapshot())
                    d = [x for x in sgd.compare(snap, oracle_snaps[r]) if x.split(":")[0] in
                         ("clauses", "literals", "h_lits_multiset", "h_full_multiset", "h_lits_ordered", "h_full_ordered")]
                    assert not d, f"round {r}: {d}"
                r += 1
            if done:
                break
        if oracle_stats is not None:
            assert r == len(oracle_stats)
        fin = s.finish()
        return to_dump(V, s.store(), fin["cnfstate"]), fin, s.rounds(), s.memory()
    finally:
        s.close()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("var", list(VARIANTS))
@pytest.mark.parametrize("name", list(SMALL))
def test_rounds_match_oracle_small(name, var):
    fam, seed, args = SMALL[name]
    flags = VARIANTS[var]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
    ed, fin, _, mem = run_engine_rounds(V, lits, offs, flags, ors, osnaps)
    assert not sgd.compare(ed, od)
    assert (ed.bits == od.bits).all() and (ed.sig == od.sig).all()
    assert mem["cuda_mallocs"] == 1   # one arena allocation per context, none in the round loop


@pytest.mark.parametrize("var", ["p5", "def"])
@pytest.mark.parametrize("name", list(MEDIUM))
def test_rounds_match_oracle_medium(name, var):
    fam, seed, args = MEDIUM[name]
    flags = VARIANTS[var]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
    ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osnaps)
    assert not sgd.compare(ed, od)


GOLD = sorted(k for k, e in SUMMARY.items() if "fingerprint" in e and "-lcvefast" not in e["flags"])
GOLD_FAST = sorted(k for k, e in SUMMARY.items() if "fingerprint" in e and "-lcvefast" in e["flags"])
FP_KEYS = ["cnfstate", "clauses", "literals", "eliminated", "forced", "resolved_words", "resolved_groups", "trail",
           "h_lits_multiset", "h_full_multiset", "h_lits_ordered", "h_full_ordered", "h_eliminated", "h_forced",
           "h_resolved_groups", "h_trail_multiset"]


@pytest.mark.parametrize("key", GOLD)
def test_engine_matches_reference_golden(key):
    """The CUDA engine against dumps of the UNMODIFIED reference GPU solver (tests/golden)."""
    e = SUMMARY[key]
    V, lits, offs = helpers.gen_cnf(e["family"], e["seed"], e["args"])
    flags = [f for f in e["flags"] if f not in ("-no-lcvefast", "-quiet")]
    ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags)
    fp, g = ed.fingerprint(), e["fingerprint"]
    diff = {k: (fp[k], g[k]) for k in FP_KEYS if fp[k] != g[k]}
    assert not diff, diff
    path = os.path.join(HERE, "golden", key + ".sgd.gz")
    if os.path.exists(path):
        ref = sgd.Dump.load(path)
        assert ref.ordered_clauses() == ed.ordered_clauses()
        assert ref.eliminated_vars() == ed.eliminated_vars()
        assert ref.resolved_groups() == ed.resolved_groups()
        assert (ref.bits == ed.bits).all() and (ref.sig == ed.sig).all()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", GOLD_FAST)
def test_engine_lcvefast_matches_reference_as_sets(key):
    """-lcvefast (the reference CLI's default election) against runs of the unmodified reference: the elected set - hence the
    eliminated variables and the clause MULTISET - is deterministic there, the order of elected[] is not (lcve.cu:204)."""
    e = SUMMARY[key]
    V, lits, offs = helpers.gen_cnf(e["family"], e["seed"], e["args"])
    flags = [f for f in e["flags"] if f not in ("-no-lcvefast", "-quiet")]
    ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags)
    fp, g = ed.fingerprint(), e["fingerprint"]
    keys = ["cnfstate", "clauses", "literals", "eliminated", "forced", "h_lits_multiset", "h_eliminated", "h_forced", "h_trail_multiset"]
    diff = {k: (fp[k], g[k]) for k in keys if fp[k] != g[k]}
    assert not diff, diff


def test_stage_prep_and_histogram():
    rng = np.random.default_rng(5)
    sizes = np.concatenate([rng.integers(1, 9, 5000), rng.integers(9, 60, 300), [1, 2, 8, 9, 250]]).astype(np.uint64)
    offs = np.zeros(len(sizes) + 1, np.uint64)
    np.cumsum(sizes, out=offs[1:])
    lits = np.empty(int(offs[-1]), np.uint32)
    for i, sz in enumerate(sizes.tolist()):  # distinct variables inside a clause
        v = rng.choice(4000, size=sz, replace=False) + 1
        lits[int(offs[i]):int(offs[i + 1])] = 2 * v + rng.integers(0, 2, sz)
    gl, gs = sigma().stage_prep(lits, offs)
    ol = lits.copy()
    os_ = np.zeros(len(sizes), np.uint32)
    helpers.oracle_lib().oracle_prep(len(sizes), ol, offs, os_)
    assert (gl == ol).all() and (gs == os_).all()
    nb = 2 * 4002
    gh = sigma().stage_histogram(lits, nb)
    oh = np.zeros(nb, np.uint32)
    helpers.oracle_lib().oracle_histogram(len(lits), lits, nb, oh)
    assert (gh == oh).all()
    assert (gh == np.bincount(lits, minlength=nb).astype(np.uint32)).all()


def test_edge_cases():
    S = sigma()
    # single clause, unit clauses, duplicate clauses, a pure literal, an immediately empty formula
    cases = [
        (3, [[2, 4, 6]]),
        (3, [[2, 4], [3, 4], [2, 5], [3, 5]]),                   # UNSAT core on x1,x2 through BVE
        (4, [[2, 4, 6], [2, 4, 6], [3, 8], [5, 8], [7, 9]]),     # duplicates + pure literals
        (2, [[2], [3, 4]]),                                      # unit in the input
        (5, [[2, 4, 6, 8, 10]] * 3),
    ]
    for V, cls in cases:
        lits = np.array([l for c in cls for l in c], np.uint32)
        offs = np.zeros(len(cls) + 1, np.uint64)
        np.cumsum([len(c) for c in cls], out=offs[1:])
        for flags in ([], ["-all"], ["--phases=1", "-no-ere"]):
            od, ors, osn = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
            ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osn)
            assert not sgd.compare(ed, od), (V, cls, flags)
    # long clauses (insertion-sort path of prep, > 8 literals) and learnt clauses in the input
    rng = np.random.default_rng(9)
    V = 300
    cls = [sorted(set((2 * (rng.choice(V, size=int(rng.integers(2, 40)), replace=False) + 1) + rng.integers(0, 2)).tolist()))
           for _ in range(900)]
    lits = np.array([l for c in cls for l in c], np.uint32)
    offs = np.zeros(len(cls) + 1, np.uint64)
    np.cumsum([len(c) for c in cls], out=offs[1:])
    meta = np.zeros(len(cls), np.uint32)
    lrn = rng.random(len(cls)) < 0.2
    meta[lrn] = 1 | (2 << 4) | (rng.integers(2, 9, int(lrn.sum())).astype(np.uint32) << 6)
    for flags, calls in (([], 1), (["-all"], 2)):
        over = helpers.opts_from_flags(flags)
        over["sigma_calls"] = calls
        od, ors, osn = helpers.run_oracle(V, lits, offs, meta=meta, snapshots=True, **over)
        s = S.Simplifier(0, flags=flags, sigma_calls=calls)
        s.load(V, lits, offs, meta=meta)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        s.close()
        assert not sgd.compare(ed, od), flags
    # bad arguments are reported, not crashed on
    s = S.Simplifier(0)
    with pytest.raises(S.SigmaError):
        s.simplify()          # nothing loaded
    s.close()


def test_reload_reuses_arena_and_is_idempotent():
    """A context can be re-run and re-loaded; the arena is allocated once while the formula fits,
    and running the simplifier twice on the same formula gives the same result."""
    S = sigma()
    V, lits, offs = helpers.gen_cnf("ksat", 12, [800, 2400, 3])
    s = S.Simplifier(0)
    s.load(V, lits, offs)
    a = s.simplify(); da = to_dump(V, s.store(), a["cnfstate"])
    b = s.simplify(); db = to_dump(V, s.store(), b["cnfstate"])
    assert not sgd.compare(da, db)
    V2, l2, o2 = helpers.gen_cnf("ksat", 14, [500, 2000, 3])   # smaller: fits the arena already held
    s.load(V2, l2, o2)
    c = s.simplify()
    od, _, _ = helpers.run_oracle(V2, l2, o2)
    assert not sgd.compare(to_dump(V2, s.store(), c["cnfstate"]), od)
    assert s.memory()["cuda_mallocs"] == 1
    s.close()


def check_model_roundtrip(V, lits, offs, store, cnfstate):
    """Every model of the simplified formula, extended over the witness stack, satisfies the
    original CNF (MODEL::extend, src/gpu/model.cpp:101-162).  The simplified formula is solved
    here by a tiny CPU routine only when it is empty (SAT by simplification) or by reusing the
    oracle's result otherwise."""
    lib = helpers.oracle_lib()
    value = np.zeros(V + 1, np.uint8)
    for u in store["trail"].tolist():
        value[u >> 1] = 0 if (u & 1) else 1
    flips = lib.oracle_extend_model(value, V, store["resolved"], len(store["resolved"]))
    return lib.oracle_check_model(value, len(offs) - 1, lits, offs), flips


def test_model_reconstruction_when_solved_by_simplification():
    # parity chains and pure-literal formulas are emptied by BVE: the witness stack alone must
    # rebuild a model of the ORIGINAL formula
    S = sigma()
    for fam, seed, args in (("parity", 41, [300]), ("parity", 43, [5000]), ("ksat", 77, [5000, 5000, 3])):
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        s = S.Simplifier(0, flags=["--phases=8"])
        s.load(V, lits, offs)
        fin = s.simplify()
        st = s.store()
        s.close()
        if fin["cnfstate"] == S.SAT:
            bad, _ = check_model_roundtrip(V, lits, offs, st, fin["cnfstate"])
            assert bad == 0, (fam, bad)
        od, _, _ = helpers.run_oracle(V, lits, offs, phases=8)
        assert od.cnfstate == fin["cnfstate"]
        assert not sgd.compare(to_dump(V, st, fin["cnfstate"]), od)


def test_full_size_properties_cfg1():
    """BASELINE config 1 at full size: parity with the oracle (it finishes in seconds) plus the
    size-independent properties - literal sortedness, signature consistency, eliminated variables
    absent from the result, idempotent histogram."""
    S = sigma()
    V, lits, offs = helpers.gen_cnf("ksat", 1, [100000, 426000, 3])
    s = S.Simplifier(0)
    s.load(V, lits, offs)
    fin = s.simplify()
    st = s.store()
    s.close()
    o = st["offs"].astype(np.int64)
    L = st["lits"]
    sz = np.diff(o)
    # sorted strictly ascending inside every clause
    inner = np.ones(len(L), bool)
    inner[o[:-1]] = False
    assert (L[1:][inner[1:]] > L[:-1][inner[1:]]).all()
    # signatures
    h = (np.uint32(1) << (L & np.uint32(31))).astype(np.uint32)
    sig = np.bitwise_or.reduceat(h, o[:-1])
    assert (sig[sz > 1] == st["sig"][sz > 1]).all()
    # eliminated variables do not occur any more
    elim = st["eliminated"]
    assert not (elim[L >> 1] & 1).any()
    assert fin["clauses"] == len(sz) and fin["literals"] == len(L)
    od, _, _ = helpers.run_oracle(V, lits, offs)
    assert not sgd.compare(to_dump(V, st, fin["cnfstate"]), od)


def test_full_size_properties_cfg2_shape():
    """BASELINE config 2 (random 5-SAT, n=1M, m=21M) at FULL size through the C ABI: properties
    only (the oracle needs minutes here) - histogram == bincount, sortedness, counters."""
    S = sigma()
    V, lits, offs = helpers.gen_cnf("ksat", 2, [1000000, 21000000, 5])
    s = S.Simplifier(0)
    s.load(V, lits, offs)
    s.begin()
    rep, done = s.round()
    hist = s.debug_hist()
    assert (hist == np.bincount(lits, minlength=len(hist)).astype(np.uint32)).all()
    while not done:
        rep, done = s.round()
    fin = s.finish()
    st = s.store()
    s.close()
    o = st["offs"].astype(np.int64)
    L = st["lits"]
    inner = np.ones(len(L), bool)
    inner[o[:-1]] = False
    assert (L[1:][inner[1:]] > L[:-1][inner[1:]]).all()
    assert fin["clauses"] == len(o) - 1 and fin["literals"] == len(L)
    assert not (st["eliminated"][L >> 1] & 1).any()
    # the eliminated variables' clauses went to the witness stack: extending any assignment that
    # satisfies the remaining clauses must satisfy the original ones.  Check the cheap direction:
    # every original clause is either still present, or contains an eliminated/forced variable.
    gone = (st["eliminated"] != 0)
    touched = np.bitwise_or.reduceat(gone[lits >> 1].astype(np.uint8), offs[:-1].astype(np.int64))
    assert int((touched == 0).sum()) <= fin["clauses"]


def sclause_stream(lits, offs, meta=None):
    """Host mirror of the reference (CNF::newClause, cnf.cuh:82-97): records {word0, sig, size, lits...} + uint64 refs."""
    n = len(offs) - 1
    sz = np.diff(offs.astype(np.int64))
    refs = (offs[:-1].astype(np.int64) + 3 * np.arange(n, dtype=np.int64)).astype(np.uint64)
    data = np.zeros(int(offs[-1]) + 3 * n, np.uint32)
    r = refs.astype(np.int64)
    data[r] = 0 if meta is None else meta
    data[r + 2] = sz
    pos = np.arange(len(lits), dtype=np.int64) + 3 * (np.repeat(np.arange(n, dtype=np.int64), sz) + 1)
    data[pos] = lits
    return data, refs


def test_load_from_reference_host_mirror():
    """sigma_load_sclauses (the SCLAUSE record stream + refs the reference copies in reflectCNF) gives the
    same result as the CSR load, with and without learnt clauses; a store_sclauses() result can be fed
    back as the input of the next call; a stream with a gap is refused."""
    S = sigma()
    rng = np.random.default_rng(3)
    for fam, seed, args, with_meta in (("ksat", 12, [800, 2400, 3], False), ("miter", 21, [30, 600, 900, 100, 8], True), ("mult", 32, [10], False)):
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        meta = None
        if with_meta:
            meta = np.zeros(len(offs) - 1, np.uint32)
            lrn = rng.random(len(meta)) < 0.15
            meta[lrn] = 1 | (1 << 4) | (rng.integers(2, 7, int(lrn.sum())).astype(np.uint32) << 6)
        calls = 2 if with_meta else 1
        a = S.Simplifier(0, sigma_calls=calls); a.load(V, lits, offs, meta=meta); ra = a.simplify(); da = to_dump(V, a.store(), ra["cnfstate"])
        data, refs = sclause_stream(lits, offs, meta)
        b = S.Simplifier(0, sigma_calls=calls); b.load_sclauses(V, data, refs); rb = b.simplify(); db = to_dump(V, b.store(), rb["cnfstate"])
        assert not sgd.compare(da, db), fam
        assert (da.bits == db.bits).all() and (da.sig == db.sig).all()
        # second call on the simplified formula, fed through the reference's own record format
        d2, r2 = a.store_sclauses()
        st = a.store()
        vstate = (st["eliminated"] != 0).astype(np.uint8)
        meta2 = (st["bits"] & np.uint32(~(2 | 4 | 8) & 0xFFFFFFFF)).astype(np.uint32)
        x = S.Simplifier(0, sigma_calls=2); x.load(V, st["lits"], st["offs"], meta=meta2, vstate=vstate); rx = x.simplify(); dx = to_dump(V, x.store(), rx["cnfstate"])
        y = S.Simplifier(0, sigma_calls=2); y.load_sclauses(V, d2, r2, vstate=vstate); ry = y.simplify(); dy = to_dump(V, y.store(), ry["cnfstate"])
        assert not sgd.compare(dx, dy), fam
        for s_ in (a, b, x, y):
            s_.close()
    V, lits, offs = helpers.gen_cnf("ksat", 12, [800, 2400, 3])
    data, refs = sclause_stream(lits, offs)
    refs[5] += 1
    s_ = S.Simplifier(0)
    with pytest.raises(S.SigmaError):
        s_.load_sclauses(V, data, refs)
    s_.close()


FULLSIZE = os.path.join(HERE, "golden", "fullsize_fingerprints.json")


FULLSIZE_KEYS = sorted(json.load(open(FULLSIZE))) if os.path.exists(FULLSIZE) else []


@pytest.mark.parametrize("name", FULLSIZE_KEYS)
def test_full_size_matches_oracle_fingerprint(name):
    """BASELINE.json configs 1-4 at FULL size: the engine's result against the fingerprint of the CPU
    oracle computed in the build container (tests/golden/make_fullsize_fingerprints.py; the oracle needs
    minutes there, the GPU box only runs the engine): same clause list in the same order with the
    same flag words, same eliminated set, witness groups and units, same per-round counters."""
    gold = json.load(open(FULLSIZE))
    if name not in gold:
        pytest.skip(f"no committed fingerprint for {name}")
    g = gold[name]
    V, lits, offs = helpers.gen_cnf(g["family"], g["seed"], g["args"])
    assert (V, len(offs) - 1, len(lits)) == (g["vars"], g["clauses_in"], g["literals_in"])
    S = sigma()
    s = S.Simplifier(0, flags=g["flags"])
    s.load(V, lits, offs)
    fin = s.simplify()
    rounds = [r for r in s.rounds() if r["kind"] == 0]
    st = s.store()
    s.close()
    got = [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds]
    assert got == g["rounds"]
    fp = to_dump(V, st, fin["cnfstate"]).fingerprint()
    diff = {k: (fp[k], v) for k, v in g["fingerprint"].items() if fp[k] != v}
    if any(f in ("-all", "-bce") for f in g["flags"]) and "h_resolved_records" in g["fingerprint"]:
        diff.pop("h_resolved_groups", None)   # blocked-clause records attach to arbitrary neighbours (sgd.py: canonical_witness); the records are compared
    assert not diff, diff


def test_oversized_bucket_path(monkeypatch):
    """k_ot_place_big (buckets too large for the shared-memory window: hot literal ranges, or more than
    2^25 literals) on small inputs: SIGMA_OT_WINDOW=0 sends every bucket through the work-unit path."""
    monkeypatch.setenv("SIGMA_OT_WINDOW", "0")
    for name in ("k3_r30", "miter_x", "mult10", "k5_r10"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True)
        ed, fin, _, _ = run_engine_rounds(V, lits, offs, [], ors, osnaps)
        assert not sgd.compare(ed, od), name
    fam, seed, args = MEDIUM["k5_20k"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True)
    ed, fin, _, _ = run_engine_rounds(V, lits, offs, [], ors, osnaps)
    assert not sgd.compare(ed, od)


def test_per_variable_inputs_vorg_vstate_assumed():
    """The per-variable inputs Solver::simplify() reads: vorg (current -> original variable
    map: the witness stack is written in original names, model.cuh), sp->vstate (inactive variables
    are never candidates, lcve.cu:87) and the incremental assumption mask (lcve.cu:88, 316-323)."""
    S = sigma()
    rng = np.random.default_rng(17)
    for name in ("k3_r30", "miter_x", "mult10", "multpar"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        vorg = np.zeros(V + 1, np.uint32)
        vorg[1:] = rng.permutation(V).astype(np.uint32) + 1 + 1000      # arbitrary original names
        used = np.zeros(V + 1, bool); used[lits >> 1] = True
        vstate = np.zeros(V + 1, np.uint8)
        vstate[~used] = 1                                                # variables without clauses: inactive, as after a host BCP
        vstate[0] = 0
        frozen_like = rng.random(V + 1) < 0.03
        vstate[frozen_like] = 3                                          # a few occurring variables marked inactive as well
        assumed = (rng.random(V + 1) < 0.05).astype(np.uint8)
        for flags in ([], ["--phases=3", "-no-ere"], ["-all"]):
            od, ors, osn = helpers.run_oracle(V, lits, offs, snapshots=True, vorg=vorg, vstate=vstate, assumed=assumed,
                                              **helpers.opts_from_flags(flags))
            s = S.Simplifier(0, flags=flags)
            s.load(V, lits, offs, vorg=vorg, vstate=vstate, assumed=assumed)
            fin = s.simplify()
            ed = to_dump(V, s.store(), fin["cnfstate"])
            rounds = [r for r in s.rounds() if r["kind"] == 0]
            s.close()
            assert not sgd.compare(ed, od), (name, flags)
            assert [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds] == \
                   [[int(x) for x in row] for row in ors], (name, flags)
            # no assumed or inactive variable was eliminated
            elim = ed.eliminated_vars()
            assert not (set(elim) & set(np.nonzero(assumed)[0].tolist())), name
            assert not (set(elim) & set(np.nonzero(vstate)[0].tolist())), name


def test_ere_queue_overflow_fallback(monkeypatch):
    """More resolvents pass the ERE filters than the queue holds: the engine sorts every list and
    searches in place (the path of the reference) - same result.  SIGMA_ERE_QUEUE_CAP=1 forces it."""
    monkeypatch.setenv("SIGMA_ERE_QUEUE_CAP", "1")
    for name in ("miter_x", "miter_a", "mult10", "k3_r42"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        for flags in ([], ["-all"]):
            od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
            ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osnaps)
            assert not sgd.compare(ed, od), (name, flags)


OPTION_SETS = [
    ["-velitsbound"],
    ["--resolventmax=6", "--phases=4"],
    ["--xormaxarity=70"],                       # beyond the small shared-memory slice: every variable runs on a full warp
    ["--xormaxarity=3", "-no-ere"],
    ["--mupos=4", "--muneg=4"],
    ["--mupos=1", "--muneg=1", "--phases=6"],   # the occurrence bound doubles per round (lcve.cu:309-311)
    ["--electionsmax=6"],
    ["--electionsmin=50"],
    ["--collectfreq=1"],
    ["--collectfreq=3", "--phases=7", "--eliminatedlitsmin=0"],
    ["--literalsmul=0.02"],                     # tight literal capacity: MEMORY_SAFE failures (bounded.cuh:266-276)
    ["--lcveclausemax=3"],
    ["--ereclausemax=4"],
    ["--eremaxoccurs=3", "--submaxoccurs=3"],
    ["-no-ve", "-no-veextend"],
    ["-no-sub", "-bce"],
    ["--phases=0"],
    ["--phases=1"],
]


@pytest.mark.parametrize("flags", OPTION_SETS, ids=lambda f: " ".join(f))
def test_option_matrix(flags):
    """The simplifier options of the reference CLI (options.cpp:24-43, options.cu:36-60), one set at a
    time, on a 3-SAT, a miter and a multiplier instance: every round equals the oracle's."""
    for name in ("k3_r42", "miter_a", "mult10", "k5_r10", "multpar", "miter_x"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
        ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osnaps)
        assert not sgd.compare(ed, od), (name, flags)


@pytest.mark.parametrize("flags", [[], ["-all"], ["--lcveclausemax=4"], ["--mupos=8", "--muneg=8", "--phases=6"], ["-no-vefunction", "--resolventmax=5"]],
                         ids=lambda f: " ".join(f) or "default")
def test_inprocessing_call_with_learnts(flags):
    """A later inprocessing call (stats.sigma.calls > 1: learnt clauses in the store, original-only
    counting in BVE, bounded.cuh / elimination.cuh `in_mode` paths) on formulas with 20 % learnt clauses."""
    S = sigma()
    rng = np.random.default_rng(23)
    for name in ("k3_r30", "miter_a", "mult10", "k4_r7"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        meta = np.zeros(len(offs) - 1, np.uint32)
        lrn = rng.random(len(meta)) < 0.2
        meta[lrn] = 1 | (rng.integers(0, 3, int(lrn.sum())).astype(np.uint32) << 4) | (rng.integers(2, 9, int(lrn.sum())).astype(np.uint32) << 6)
        over = helpers.opts_from_flags(flags)
        over["sigma_calls"] = 3
        od, ors, osn = helpers.run_oracle(V, lits, offs, meta=meta, snapshots=True, **over)
        s = S.Simplifier(0, flags=flags, sigma_calls=3)
        s.load(V, lits, offs, meta=meta)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        rounds = [r for r in s.rounds() if r["kind"] == 0]
        s.close()
        assert not sgd.compare(ed, od), (name, flags)
        assert (ed.bits == od.bits).all() and (ed.sig == od.sig).all(), (name, flags)
        assert [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds] == \
               [[int(x) for x in row] for row in ors], (name, flags)


FUZZ_FLAGS = [[], ["-all"], ["-bce"], ["--lcveclausemax=5"], ["--mupos=6", "--muneg=6"], ["--phases=7", "--eliminatedlitsmin=0"],
              ["-no-vefunction"], ["--resolventmax=7"], ["--xormaxarity=4"], ["--collectfreq=1"], ["-velitsbound"],
              ["--electionsmax=8"], ["--literalsmul=0.1"], ["-no-sub"], ["-no-ere", "--phases=3"], ["--ereclausemax=6"],
              ["-aggresivesort"], ["--lcveclausemax=3", "--phases=6"], ["--xormaxarity=70"]]


def random_cnf(rng, V, C, kmin, kmax, dup=0.02):
    """Mixed clause sizes, no tautologies, a few duplicate clauses, a few units."""
    cls = []
    for _ in range(C):
        k = int(rng.integers(kmin, kmax + 1))
        vs = rng.choice(V, size=min(k, V), replace=False) + 1
        cls.append((2 * vs + rng.integers(0, 2, len(vs))).astype(np.uint32))
    for _ in range(int(C * dup)):
        cls.append(cls[int(rng.integers(0, len(cls)))].copy())
    order = rng.permutation(len(cls))
    cls = [cls[i] for i in order]
    lits = np.concatenate(cls).astype(np.uint32)
    offs = np.zeros(len(cls) + 1, np.uint64)
    np.cumsum([len(c) for c in cls], out=offs[1:])
    return lits, offs


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_FUZZ_SEEDS", "64")))))
def test_random_fuzz(seed):
    """Random small formulas (mixed clause sizes 2-12, duplicates, optional learnt clauses, inactive and
    assumed variables) under a random option set: every round and the final result equal the oracle's."""
    S = sigma()
    rng = np.random.default_rng(1000 + seed)
    V = int(rng.integers(30, 500))
    ratio = float(rng.choice([1.5, 2.5, 4.0, 6.0, 10.0]))
    kmin = int(rng.integers(2, 4)); kmax = int(rng.integers(kmin, 13))
    lits, offs = random_cnf(rng, V, max(8, int(V * ratio)), kmin, kmax)
    flags = list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    if rng.random() < 0.3:
        flags += list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    calls = int(rng.integers(1, 4))
    meta = None
    if calls > 1 and rng.random() < 0.7:
        meta = np.zeros(len(offs) - 1, np.uint32)
        sz = np.diff(offs.astype(np.int64))
        lrn = (rng.random(len(meta)) < 0.25) & (sz > 1)
        meta[lrn] = 1 | (rng.integers(0, 3, int(lrn.sum())).astype(np.uint32) << 4) | (rng.integers(2, 9, int(lrn.sum())).astype(np.uint32) << 6)
    vstate = assumed = vorg = None
    if rng.random() < 0.4:
        vstate = (rng.random(V + 1) < 0.04).astype(np.uint8) * 3; vstate[0] = 0
    if rng.random() < 0.4:
        assumed = (rng.random(V + 1) < 0.06).astype(np.uint8)
    if seed >= 64 and rng.random() < 0.3:      # (seeds < 64 keep the formulas of the first sweeps)
        vorg = np.zeros(V + 1, np.uint32); vorg[1:] = rng.permutation(V).astype(np.uint32) + 1 + int(rng.integers(0, 5000))
    if seed >= 64 and rng.random() < 0.15:     # a few unit clauses in the input (the host normally propagates them first)
        nu = int(rng.integers(1, 4))
        uv = rng.choice(V, size=nu, replace=False) + 1
        ul = (2 * uv + rng.integers(0, 2, nu)).astype(np.uint32)
        lits = np.concatenate([lits, ul]).astype(np.uint32)
        offs = np.concatenate([offs, offs[-1] + np.arange(1, nu + 1, dtype=np.uint64)]).astype(np.uint64)
        if meta is not None:
            meta = np.concatenate([meta, np.zeros(nu, np.uint32)])
    try:
        over = helpers.opts_from_flags(flags)
    except KeyError:
        pytest.skip("flag combination not expressible")
    over["sigma_calls"] = calls
    od, ors, _ = helpers.run_oracle(V, lits, offs, meta=meta, vorg=vorg, vstate=vstate, assumed=assumed, **over)
    s = S.Simplifier(0, flags=flags, sigma_calls=calls)
    try:
        s.load(V, lits, offs, meta=meta, vorg=vorg, vstate=vstate, assumed=assumed)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        rounds = [r for r in s.rounds() if r["kind"] == 0]
    finally:
        s.close()
    ctx = (seed, V, len(offs) - 1, kmin, kmax, flags, calls)
    assert od.cnfstate == fin["cnfstate"], ctx
    assert [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds] == \
           [[int(x) for x in row] for row in ors], ctx
    assert not sgd.compare(ed, od), ctx


def skewed_cnf(rng, V, C, kmax, hubs, hub_p):
    """Hub variables (occurrence lists of hundreds to thousands of entries: every list-sort class, the
    32-lane groups, electionsmax / submaxoccurs bounds) and clause sizes up to kmax (> 8: the long-clause
    paths of prep, partition and the MIS walk)."""
    hub = rng.choice(V, size=hubs, replace=False) + 1
    cls = []
    for _ in range(C):
        k = int(min(V - 1, max(2, rng.geometric(0.35) + 1 if rng.random() < 0.8 else rng.integers(9, kmax + 1))))
        vs = set((rng.choice(V, size=k, replace=False) + 1).tolist())
        if rng.random() < hub_p:
            vs.add(int(hub[int(rng.integers(0, hubs))]))
        vs = np.array(sorted(vs), np.uint32)
        cls.append((2 * vs + rng.integers(0, 2, len(vs))).astype(np.uint32))
    lits = np.concatenate(cls).astype(np.uint32)
    offs = np.zeros(len(cls) + 1, np.uint64)
    np.cumsum([len(c) for c in cls], out=offs[1:])
    return lits, offs


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_FUZZ2_SEEDS", "24")))))
def test_random_fuzz_skewed(seed):
    S = sigma()
    rng = np.random.default_rng(5000 + seed)
    V = int(rng.integers(300, 4000))
    C = int(V * float(rng.choice([2.0, 3.5, 5.0])))
    lits, offs = skewed_cnf(rng, V, C, int(rng.integers(12, 60)), int(rng.integers(1, 6)), float(rng.choice([0.05, 0.3, 0.8])))
    flags = list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    if rng.random() < 0.5:
        flags += [str(rng.choice(["--electionsmax=40", "--submaxoccurs=20", "--eremaxoccurs=30", "--mupos=256", "--bcemaxoccurs=25"]))]
        if flags[-1] == "--mupos=256":
            flags.append("--muneg=256")
    calls = int(rng.integers(1, 3))
    over = helpers.opts_from_flags(flags); over["sigma_calls"] = calls
    od, ors, _ = helpers.run_oracle(V, lits, offs, **over)
    s = S.Simplifier(0, flags=flags, sigma_calls=calls)
    try:
        s.load(V, lits, offs)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        rounds = [r for r in s.rounds() if r["kind"] == 0]
    finally:
        s.close()
    ctx = (seed, V, C, flags, calls)
    assert od.cnfstate == fin["cnfstate"], ctx
    if od.cnfstate != 0:
        assert [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds] == \
               [[int(x) for x in row] for row in ors], ctx
    assert not sgd.compare(ed, od), ctx


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_FUZZ3_SEEDS", "24")))))
def test_random_fuzz_structured(seed):
    """Gate-rich formulas (Tseitin miters with a random AND/XOR mix and rewriting rate, array multipliers,
    parity chains) of random size under random options: equivalence / AND-OR / ITE / XOR / function-table
    substitution paths of BVE, SUB on gate clauses, ERE on redundant resolvents."""
    S = sigma()
    rng = np.random.default_rng(9000 + seed)
    kind = int(rng.integers(0, 4))
    if kind == 0:
        fam, args = "miter", [int(rng.integers(8, 80)), int(rng.integers(60, 2500)), int(rng.integers(0, 1001)), int(rng.integers(0, 400)), int(rng.integers(1, 33))]
    elif kind == 1:
        fam, args = "mult", [int(rng.integers(3, 15))]
    elif kind == 2:
        fam, args = "parity", [int(rng.integers(10, 800))]
    else:
        fam, args = "multpar", [int(rng.integers(3, 9)), int(rng.integers(10, 400))]
    V, lits, offs = helpers.gen_cnf(fam, 7000 + seed, args)
    flags = list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    if rng.random() < 0.4:
        flags += list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    calls = int(rng.integers(1, 3))
    over = helpers.opts_from_flags(flags); over["sigma_calls"] = calls
    od, ors, osn = helpers.run_oracle(V, lits, offs, snapshots=True, **over)
    ed, fin, _, _ = (None, None, None, None)
    s = S.Simplifier(0, flags=flags, sigma_calls=calls)
    try:
        s.load(V, lits, offs)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        rounds = [r for r in s.rounds() if r["kind"] == 0]
    finally:
        s.close()
    ctx = (seed, fam, args, flags, calls)
    assert od.cnfstate == fin["cnfstate"], ctx
    if od.cnfstate != 0:
        assert [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds] == \
               [[int(x) for x in row] for row in ors], ctx
    assert not sgd.compare(ed, od), ctx


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_FUZZ4_SEEDS", "10")))))
def test_random_fuzz_medium(seed):
    """Medium instances (20 k - 150 k variables): several partition tiles and buckets, staged and direct
    partition tiles, chunked (dense) and clause-pass (sparse) election, all list-sort classes."""
    S = sigma()
    rng = np.random.default_rng(12000 + seed)
    kind = int(rng.integers(0, 5))
    if kind <= 2:
        k = int(rng.choice([3, 4, 5, 7]))
        n = int(rng.integers(20000, 150000))
        ratio = {3: 4.26, 4: 9.0, 5: float(rng.choice([12.0, 21.0])), 7: 40.0}[k] * float(rng.choice([0.6, 1.0]))
        fam, args = "ksat", [n, int(n * ratio), k]
    elif kind == 3:
        fam, args = "miter", [int(rng.integers(200, 1500)), int(rng.integers(20000, 120000)), int(rng.integers(0, 1001)), int(rng.integers(0, 300)), 32]
    else:
        fam, args = "multpar", [int(rng.integers(24, 64)), int(rng.integers(5000, 60000))]
    V, lits, offs = helpers.gen_cnf(fam, 300 + seed, args)
    flags = list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    over = helpers.opts_from_flags(flags)
    od, ors, _ = helpers.run_oracle(V, lits, offs, **over)
    s = S.Simplifier(0, flags=flags)
    try:
        s.load(V, lits, offs)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        rounds = [r for r in s.rounds() if r["kind"] == 0]
    finally:
        s.close()
    ctx = (seed, fam, args, flags)
    assert od.cnfstate == fin["cnfstate"], ctx
    if od.cnfstate != 0:
        assert [[r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]] for r in rounds] == \
               [[int(x) for x in row] for row in ors], ctx
    assert not sgd.compare(ed, od), ctx


@pytest.mark.parametrize("flags", [["-aggresivesort"], ["-aggresivesort", "-all"], ["-aggresivesort", "--phases=2", "-no-ere"]], ids=lambda f: " ".join(f))
def test_aggressive_write_back_order(flags):
    """-aggresivesort (cacheCNF, cnf.cu:232-233): the clauses leave sorted by OLIST_CMP (size, first literal,
    last literal, signature, ref); per-round snapshots stay in ref order.  CSR and SCLAUSE-stream outputs."""
    S = sigma()
    for name in ("k3_r30", "miter_x", "mult10", "k5_r10", "multpar"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
        ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osnaps)
        assert not sgd.compare(ed, od), (name, flags)
        assert ed.ordered_clauses() == od.ordered_clauses()
        oc = ed.ordered_clauses()
        key = [(len(c), c[0], c[-1]) for c in oc]
        assert key == sorted(key), name                      # the order really is the OLIST_CMP order
        s = S.Simplifier(0, flags=flags)
        s.load(V, lits, offs)
        s.simplify()
        data, refs = s.store_sclauses()
        s.close()
        rec = [tuple(data[int(r) + 3:int(r) + 3 + int(data[int(r) + 2])].tolist()) for r in refs]
        assert rec == oc, name
