"""CPU checks of the oracle's device-DRAT restatement (src/gpu/proof.cu, proofutils.cuh).

The reference holds no golden proof for this path and its GPU binary cannot run here, so the stream is
pinned semantically: every added line must follow from the formula by unit propagation (RUP), every
deleted line must name a clause that exists, every clause of the simplified result must be in the
checker's database at the end, and switching the proof on must not change the simplification."""
import numpy as np
import pytest

import helpers
import sgd

CASES = {
    "k3_r30": ("ksat", 12, [800, 2400, 3]),
    "k3_r42": ("ksat", 11, [600, 2520, 3]),
    "miter_a": ("miter", 22, [40, 700, 300, 200, 8]),
    "mult6": ("mult", 31, [6]),
    "mult10": ("mult", 32, [10]),
    "parity": ("parity", 41, [300]),
    "multpar": ("multpar", 51, [5, 120]),
}
FLAGSETS = {"def": [], "all": ["-all"], "p2_bce": ["--phases=2", "-bce"], "nofun": ["-no-vefunction"]}


def clauses_of(lits, offs, vorg=None):
    out = []
    for i in range(len(offs) - 1):
        c = [int(x) for x in lits[int(offs[i]):int(offs[i + 1])]]
        if vorg is not None:
            c = [(int(vorg[l >> 1]) << 1) | (l & 1) for l in c]
        out.append(tuple(c))
    return out


def check_stream(chunks, original, result, defer_deletions=False):
    """forward check of the chunks; returns (#added, #deleted)"""
    ck = helpers.RupChecker(original)
    na = nd = 0
    for r, ch in enumerate(chunks):
        pending = []
        for kind, l in helpers.drat_parse(ch):
            if kind == b"a":
                assert ck.rup(l), f"chunk {r}: added clause {l} is not RUP"
                ck.add(l)
                na += 1
            elif defer_deletions:
                pending.append(l)
            else:
                assert ck.delete(l), f"chunk {r}: deleted clause {l} is not in the formula"
                nd += 1
        for l in pending:
            assert ck.delete(l), f"chunk {r}: deleted clause {l} is not in the formula"
            nd += 1
    # a clause of the result is an original or added clause, or one of those shortened by root-level units in prop()
    # (bcp_apply_k, elimbcp.cu:124-141, writes no proof line: the shortened clause follows by unit propagation)
    for c in result:
        assert ck.has(c) or ck.rup(c), f"result clause {c} does not follow from the formula and the proof"
    return na, nd


@pytest.mark.parametrize("fl", list(FLAGSETS))
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_proof_is_valid_drat(name, fl):
    fam, seed, args = CASES[name]
    flags = FLAGSETS[fl]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    plain, rs0, _ = helpers.run_oracle(V, lits, offs, **helpers.opts_from_flags(flags))
    d, rs, _ = helpers.run_oracle(V, lits, offs, proof=True, **helpers.opts_from_flags(flags))
    assert not sgd.compare(d, plain) and (rs == rs0).all()      # the proof guards never fire at these sizes
    chunks = d.extra["proof"]
    assert all(len(c) <= d.extra["proof_cap"] for c in chunks)   # cuPROOF::alloc capacity (simplify.cu:128-132)
    na, nd = check_stream(chunks, clauses_of(lits, offs), clauses_of(d.lits, d.offs))
    if name != "parity" and len(rs):
        assert na > 0


def test_oracle_proof_uses_original_literals():
    """vorg maps the working variables to sparse original numbers: multi-byte varints, ORIGINIZELIT."""
    fam, seed, args = CASES["mult10"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    rng = np.random.default_rng(5)
    vorg = np.zeros(V + 1, np.uint32)
    vorg[1:] = np.sort(rng.choice(np.arange(1, 3_000_000, dtype=np.uint32), V, replace=False))   # up to 4-byte literals
    d, rs, _ = helpers.run_oracle(V, lits, offs, vorg=vorg, proof=True)
    ident, _, _ = helpers.run_oracle(V, lits, offs, proof=True)
    # same lines, renamed
    for a, b in zip(d.extra["proof"], ident.extra["proof"]):
        la, lb = helpers.drat_parse(a), helpers.drat_parse(b)
        assert [(k, tuple((int(vorg[x >> 1]) << 1) | (x & 1) for x in l)) for k, l in lb] == la
    assert d.extra["proof_cap"] > ident.extra["proof_cap"]
    check_stream(d.extra["proof"], clauses_of(lits, offs, vorg), clauses_of(d.lits, d.offs, vorg))


def test_drat_varint_roundtrip():
    lits = [2, 3, 127, 128, 255, 16383, 16384, 2_097_151, 2_097_152, 268_435_455, 268_435_456, 0x7FFFFFFF]
    raw = bytearray(b"a")
    for l in lits:
        v = l
        while v & ~0x7F:
            raw.append((v & 0x7F) | 0x80)
            v >>= 7
        raw.append(v)
    raw.append(0)
    assert helpers.drat_parse(bytes(raw)) == [(b"a", tuple(lits))]


def test_rup_checker_rejects_a_wrong_line():
    ck = helpers.RupChecker([(2, 4), (3, 6)])     # (x1 v x2), (-x1 v x3)
    assert ck.rup((4, 6))                          # resolvent
    assert not ck.rup((4,))                        # x2 alone does not follow
    assert not ck.delete((2, 6))


# ---------------------------------------------------------------------------------------------
# proof files written by the UNMODIFIED reference GPU binary on a B200 (tests/golden/make_golden_proofs.py)
import glob  # noqa: E402
import gzip  # noqa: E402
import json  # noqa: E402
import os  # noqa: E402

PROOF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proof")
PROOF_SUMMARY = json.load(open(os.path.join(PROOF_DIR, "summary.json")))


@pytest.mark.parametrize("key", sorted(PROOF_SUMMARY))
def test_oracle_proof_matches_reference_file(key):
    """The reference's proof file after `ref_driver <cnf> -no-lcvefast -proof` holds exactly the lines of the oracle's
    chunks (with -no-solve-like driving no host line is written on these instances); compared as multisets, since
    the reference's threads reserve their lines in scheduling order."""
    e = PROOF_SUMMARY[key]
    raw = gzip.open(os.path.join(PROOF_DIR, key + ".drat.gz")).read()
    assert len(raw) == e["proof_bytes"]
    V, lits, offs = helpers.gen_cnf(e["family"], e["seed"], e["args"])
    flags = [f for f in e["flags"] if f != "-proof"]
    d, _, _ = helpers.run_oracle(V, lits, offs, proof=True, **helpers.opts_from_flags(flags))
    # the simplified CNF of the same run (ERE runs carry --ereminthreads=32, see tests/golden/make_golden.py ERE_LAUNCH_FIX)
    if "fingerprint" in e and ("-no-ere" in flags or "--ereminthreads=32" in flags):
        fp, g = d.fingerprint(), e["fingerprint"]
        keys = ["cnfstate", "clauses", "literals", "eliminated", "h_full_ordered", "h_eliminated", "h_resolved_groups", "h_trail_multiset"]
        assert {k: (fp[k], g[k]) for k in keys if fp[k] != g[k]} == {}
    mine = b"".join(d.extra["proof"])
    assert len(mine) == len(raw)
    assert helpers.drat_canonical(mine) == helpers.drat_canonical(raw)


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("SIGMA_PROOF_CPU_SEEDS", "48")))))
def test_oracle_proof_fuzz(seed):
    """the generator of test_zz_gpu_proof.py::test_proof_stream_fuzz_random: random formulas with learnt clauses,
    inactive / assumed variables, sparse original numbering and units, under random option sets"""
    from test_gpu_parity import FUZZ_FLAGS, random_cnf
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
    if rng.random() < 0.5:
        vorg = np.zeros(V + 1, np.uint32)
        vorg[1:] = rng.permutation(V).astype(np.uint32) * int(rng.choice([1, 40, 3000])) + 1 + int(rng.integers(0, 5000))
    try:
        over = helpers.opts_from_flags(flags)
    except KeyError:
        pytest.skip("flag combination not expressible")
    over["sigma_calls"] = calls
    d, _, _ = helpers.run_oracle(V, lits, offs, meta=meta, vorg=vorg, vstate=vstate, assumed=assumed, proof=True, **over)
    if d.cnfstate == 0:
        pytest.skip("UNSAT by propagation")
    assert all(len(c) <= d.extra["proof_cap"] for c in d.extra["proof"])
    check_stream(d.extra["proof"], clauses_of(lits, offs, vorg), clauses_of(d.lits, d.offs, vorg))


def test_binary_to_text_drat():
    """tools/drat_text.py against the reference's own proof file of the parity instance and a hand-made stream"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(PROOF_DIR), "..", "..", "tools"))
    import drat_text
    assert drat_text.to_text(b"a\x04\x07\x00d\x80\x02\x00") == "2 -3 0\nd 128 0\n"
    raw = gzip.open(os.path.join(PROOF_DIR, "k3_r42__nofun.drat.gz")).read()
    txt = drat_text.to_text(raw).splitlines()
    parsed = helpers.drat_parse(raw)
    assert len(txt) == len(parsed) == 71
    for line, (kind, lits) in zip(txt, parsed):
        want = [(-(l >> 1) if l & 1 else l >> 1) for l in lits]
        assert line == ("d " if kind == b"d" else "") + " ".join(map(str, want)) + " 0"
    with pytest.raises(ValueError):
        drat_text.to_text(b"x\x04\x00")
