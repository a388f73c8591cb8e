"""GPU tests of the device DRAT proof stream (opts.proof_en / flag -proof; include/sigma.h sigma_proof_*),
the SURVEY 8(f)-2 row: the engine's stream against the oracle's restatement of src/gpu/proof.cu +
proofutils.cuh, chunk by chunk (one chunk per cacheProof/writeProof point), as multisets of lines - the
reference itself appends the lines of different variables in whatever order their threads reserve space -
and, in the engine's own order, through a forward RUP check."""
import numpy as np
import pytest

import helpers
import sgd
from test_oracle_proof import CASES, check_stream, clauses_of

pytestmark = pytest.mark.gpu

FLAGSETS = {"def": [], "all": ["-all"], "p2_bce": ["--phases=2", "-bce"], "nofun_noere": ["-no-vefunction", "-no-ere"]}


def sigma():
    from parafrost_b200 import sigma as s
    return s


def to_dump(V, st, state=2):
    return sgd.Dump.from_arrays(V, state, st["bits"], st["sig"], st["offs"], st["lits"], st["eliminated"], st["resolved"], st["trail"])


def run_engine(V, lits, offs, flags, vorg=None, meta=None, sink=None, **opts):
    s = sigma().Simplifier(0, flags=list(flags) + ["-proof"], **opts)
    try:
        if sink is not None:
            s.set_proof_sink(sink)
        s.load(V, lits, offs, vorg=vorg, meta=meta)
        fin = s.simplify()
        chunks, cap = s.proof_chunks()
        return to_dump(V, s.store(), fin["cnfstate"]), chunks, cap, fin
    finally:
        s.close()


@pytest.mark.parametrize("fl", list(FLAGSETS))
@pytest.mark.parametrize("name", list(CASES))
def test_proof_stream_matches_oracle(name, fl):
    fam, seed, args = CASES[name]
    flags = FLAGSETS[fl]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, _ = helpers.run_oracle(V, lits, offs, proof=True, **helpers.opts_from_flags(flags))
    got = []
    ed, chunks, cap, fin = run_engine(V, lits, offs, flags, sink=got.append)
    assert not sgd.compare(ed, od)                       # the proof passes do not disturb the simplification
    assert (ed.bits == od.bits).all() and (ed.sig == od.sig).all()
    assert cap == od.extra["proof_cap"]                  # cuPROOF::count x 1.5 (simplify.cu:128-132)
    ochunks = od.extra["proof"]
    assert [len(c) for c in chunks] == [len(c) for c in ochunks]
    for r, (a, b) in enumerate(zip(chunks, ochunks)):
        assert helpers.drat_canonical(a) == helpers.drat_canonical(b), f"chunk {r}"
    assert b"".join(got) == b"".join(chunks)             # the sink saw every non-empty chunk, in order
    # the engine's own line order is a valid derivation (deletions of a round applied at its end)
    check_stream(chunks, clauses_of(lits, offs), clauses_of(ed.lits, ed.offs), defer_deletions=True)


def test_proof_stream_original_numbering_and_round_api():
    """sparse vorg (multi-byte varints) + sigma_begin/round/finish instead of sigma_run"""
    fam, seed, args = CASES["mult10"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    rng = np.random.default_rng(5)
    vorg = np.zeros(V + 1, np.uint32)
    vorg[1:] = np.sort(rng.choice(np.arange(1, 3_000_000, dtype=np.uint32), V, replace=False))
    od, _, _ = helpers.run_oracle(V, lits, offs, vorg=vorg, proof=True)
    s = sigma().Simplifier(0, flags=["-proof"])
    try:
        s.load(V, lits, offs, vorg=vorg)
        for _ in range(2):   # a second run on the same load starts a fresh stream
            s.begin()
            while True:
                _, done = s.round()
                if done:
                    break
            s.finish()
            chunks, cap = s.proof_chunks()
            assert cap == od.extra["proof_cap"]
            assert [helpers.drat_canonical(c) for c in chunks] == [helpers.drat_canonical(c) for c in od.extra["proof"]]
        st = s.store()
    finally:
        s.close()
    check_stream(chunks, clauses_of(lits, offs, vorg), clauses_of(st["lits"], st["offs"], vorg), defer_deletions=True)


def test_proof_stream_with_learnt_clauses():
    """later inprocessing call: learnt clauses in the input (sigma_calls > 1), SUB may delete or strengthen them"""
    fam, seed, args = CASES["k3_r30"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    rng = np.random.default_rng(11)
    C = len(offs) - 1
    meta = np.zeros(C, np.uint32)
    learnt = rng.random(C) < 0.15
    meta[learnt] = 1 | (2 << 4) | (rng.integers(2, 9, learnt.sum()).astype(np.uint32) << 6)
    od, _, _ = helpers.run_oracle(V, lits, offs, meta=meta, proof=True, sigma_calls=2)
    ed, chunks, cap, _ = run_engine(V, lits, offs, [], meta=meta, sigma_calls=2)
    assert not sgd.compare(ed, od)
    assert [helpers.drat_canonical(c) for c in chunks] == [helpers.drat_canonical(c) for c in od.extra["proof"]]


def test_proof_off_leaves_no_stream_and_late_enable_is_refused():
    fam, seed, args = CASES["mult6"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    s = sigma().Simplifier(0)
    try:
        s.load(V, lits, offs)
        s.simplify()
        chunks, _ = s.proof_chunks()
        assert chunks == []
        s.optSimp(flags=["-proof"])          # the stream buffer is carved by sigma_load
        with pytest.raises(sigma().SigmaError):
            s.simplify()
        s.load(V, lits, offs)                # reload with the option set: now it works
        s.simplify()
        chunks, _ = s.proof_chunks()
        assert sum(len(c) for c in chunks) > 0
    finally:
        s.close()


# ---------------------------------------------------------------------------------------------
# randomised: the proof stream under random formulas and option sets
import os  # noqa: E402

from test_gpu_parity import FUZZ_FLAGS, random_cnf  # noqa: E402


def _compare_streams(V, lits, offs, flags, calls, ctx, meta=None, vorg=None, vstate=None, assumed=None):
    try:
        over = helpers.opts_from_flags(flags)
    except KeyError:
        pytest.skip("flag combination not expressible")
    over["sigma_calls"] = calls
    od, ors, _ = helpers.run_oracle(V, lits, offs, meta=meta, vorg=vorg, vstate=vstate, assumed=assumed, proof=True, **over)
    s = sigma().Simplifier(0, flags=list(flags) + ["-proof"], sigma_calls=calls)
    try:
        s.load(V, lits, offs, meta=meta, vorg=vorg, vstate=vstate, assumed=assumed)
        fin = s.simplify()
        ed = to_dump(V, s.store(), fin["cnfstate"])
        chunks, cap = s.proof_chunks()
    finally:
        s.close()
    assert od.cnfstate == fin["cnfstate"], ctx
    if od.cnfstate == 0:
        return   # UNSAT by propagation: the answer is the result (the conflict surfaces on a schedule-dependent variable)
    assert not sgd.compare(ed, od), ctx
    assert cap == od.extra["proof_cap"], ctx
    assert [len(c) for c in chunks] == [len(c) for c in od.extra["proof"]], ctx
    for r, (a, b) in enumerate(zip(chunks, od.extra["proof"])):
        assert helpers.drat_canonical(a) == helpers.drat_canonical(b), (ctx, r)


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_PROOF_FUZZ_SEEDS", "32")))))
def test_proof_stream_fuzz_random(seed):
    rng = np.random.default_rng(31000 + seed)
    V = int(rng.integers(30, 400))
    ratio = float(rng.choice([1.5, 2.5, 4.0, 6.0]))
    kmin = int(rng.integers(2, 4)); kmax = int(rng.integers(kmin, 10))
    lits, offs = random_cnf(rng, V, max(8, int(V * ratio)), kmin, kmax)
    flags = list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    if rng.random() < 0.3:
        flags += list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    calls = int(rng.integers(1, 4))
    meta = vorg = vstate = assumed = None
    if calls > 1 and rng.random() < 0.7:
        meta = np.zeros(len(offs) - 1, np.uint32)
        sz = np.diff(offs.astype(np.int64))
        lrn = (rng.random(len(meta)) < 0.25) & (sz > 1)
        meta[lrn] = 1 | (rng.integers(0, 3, int(lrn.sum())).astype(np.uint32) << 4) | (rng.integers(2, 9, int(lrn.sum())).astype(np.uint32) << 6)
    if rng.random() < 0.3:
        vstate = (rng.random(V + 1) < 0.04).astype(np.uint8) * 3; vstate[0] = 0
    if rng.random() < 0.3:
        assumed = (rng.random(V + 1) < 0.06).astype(np.uint8)
    if rng.random() < 0.5:   # sparse original numbering: literals of 1 to 4 proof bytes
        vorg = np.zeros(V + 1, np.uint32)
        vorg[1:] = rng.permutation(V).astype(np.uint32) * int(rng.choice([1, 40, 3000])) + 1 + int(rng.integers(0, 5000))
    _compare_streams(V, lits, offs, flags, calls, (seed, V, len(offs) - 1, kmin, kmax, flags, calls), meta, vorg, vstate, assumed)


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_PROOF_FUZZ2_SEEDS", "24")))))
def test_proof_stream_fuzz_structured(seed):
    rng = np.random.default_rng(32000 + seed)
    kind = int(rng.integers(0, 4))
    if kind == 0:
        fam, args = "miter", [int(rng.integers(8, 60)), int(rng.integers(60, 1500)), int(rng.integers(0, 1001)), int(rng.integers(0, 400)), int(rng.integers(1, 33))]
    elif kind == 1:
        fam, args = "mult", [int(rng.integers(3, 13))]
    elif kind == 2:
        fam, args = "parity", [int(rng.integers(10, 600))]
    else:
        fam, args = "multpar", [int(rng.integers(3, 9)), int(rng.integers(10, 300))]
    V, lits, offs = helpers.gen_cnf(fam, 8000 + seed, args)
    flags = list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    if rng.random() < 0.4:
        flags += list(FUZZ_FLAGS[int(rng.integers(0, len(FUZZ_FLAGS)))])
    calls = int(rng.integers(1, 3))
    _compare_streams(V, lits, offs, flags, calls, (seed, fam, args, flags, calls))
