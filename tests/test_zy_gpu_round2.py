"""Round-2 additions, all through the C ABI on a GPU: the -lcvefast election (filtered-candidate MIS, lcve.cu:150-217),
the device-resident continuation (sigma_continue / sigma_device_view, simplify.cu:221-229, solver.hpp:694-705), the compact
store, the per-prop trail ranges and the capacity guard."""
import ctypes as C

import numpy as np
import pytest

import helpers
import sgd
from test_gpu_parity import SMALL, MEDIUM, random_cnf, run_engine_rounds, sigma, to_dump

pytestmark = pytest.mark.gpu

FAST_SETS = [["-lcvefast"], ["-lcvefast", "-no-vefunction"], ["-lcvefast", "-all"], ["-lcvefast", "--mupos=6", "--muneg=6"],
             ["-lcvefast", "--lcveclausemax=3"], ["-lcvefast", "--electionsmax=8", "--phases=6"]]


@pytest.mark.parametrize("flags", FAST_SETS, ids=lambda f: " ".join(f))
def test_lcvefast_rounds_match_oracle(flags):
    """Every round of the -lcvefast mode equals the oracle's restatement of mis_init_k / mis_round_k / mis_freeze_k."""
    for name in SMALL:
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
        ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osnaps)
        assert not sgd.compare(ed, od), (name, flags)


def test_lcvefast_elects_a_superset_schedule_on_medium():
    fam, seed, args = MEDIUM["miter_50k"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, _ = helpers.run_oracle(V, lits, offs, **helpers.opts_from_flags(["-lcvefast"]))
    ed, fin, rounds, _ = run_engine_rounds(V, lits, offs, ["-lcvefast"], ors, None)
    assert not sgd.compare(ed, od)
    _, ors0, _ = helpers.run_oracle(V, lits, offs)
    assert int(ors[0][0]) >= int(ors0[0][0])      # stops are filters: the first round elects at least as many variables


@pytest.mark.parametrize("seed", list(range(24)))
def test_lcvefast_fuzz(seed):
    rng = np.random.default_rng(52000 + seed)
    V = int(rng.integers(30, 500))
    lits, offs = random_cnf(rng, V, max(8, int(V * float(rng.choice([1.5, 3.0, 4.2, 6.0])))), 2, int(rng.integers(3, 9)))
    flags = ["-lcvefast"] + list([[], ["-all"], ["--lcveclausemax=4"], ["--mupos=4", "--muneg=4"], ["-no-vefunction"], ["--electionsmax=6"]][int(rng.integers(0, 6))])
    od, ors, osn = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
    if od.cnfstate == 0:
        s = sigma().Simplifier(0, flags=flags)
        s.load(V, lits, offs)
        assert s.simplify()["cnfstate"] == 0
        s.close()
        return
    ed, fin, _, _ = run_engine_rounds(V, lits, offs, flags, ors, osn)
    assert not sgd.compare(ed, od), flags


def second_call_inputs(V, d, extra=None):
    """What the host hands to the next inprocessing call after `d` (a Dump of the previous result): the clause list with its
    learnt bits, inactive = eliminated (not forced) or on the trail."""
    lits, offs, meta = d.lits.copy(), d.offs.copy(), (d.bits & ~np.uint32(2 | 4 | 8)).astype(np.uint32)
    meta[(d.bits & 1) == 0] = 0
    vstate = np.zeros(V + 1, np.uint8)
    el = np.asarray(d.eliminated[: V + 1])
    vstate[(el != 0) & ((el & 4) == 0)] = 3
    for u in d.trail.tolist():
        vstate[u >> 1] = 2
    if extra is not None:
        xl, xo, xm = extra
        lits = np.concatenate([lits, xl]); offs = np.concatenate([offs, xo[1:] + offs[-1]]); meta = np.concatenate([meta, xm])
    return lits, offs.astype(np.uint64), meta, vstate


@pytest.mark.parametrize("name", ["k3_r30", "miter_a", "mult10", "multpar", "k4_r7"])
@pytest.mark.parametrize("delta", [False, True])
def test_resident_second_call_matches_oracle(name, delta):
    """Two inprocessing calls, the second one on the RESIDENT result (sigma_continue, no clause crosses PCIe again except
    the delta): equal to the oracle run twice, the second time with sigma_calls = 2 on the first result."""
    fam, seed, args = SMALL[name]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    flags = ["--phases=2", "-no-ere"]
    od1, _, _ = helpers.run_oracle(V, lits, offs, **helpers.opts_from_flags(flags))
    rng = np.random.default_rng(7)
    extra = None
    if delta:   # a few learnt clauses over variables still active
        act = np.array([v for v in range(1, V + 1) if od1.eliminated[v] == 0 and v not in {u >> 1 for u in od1.trail.tolist()}], np.uint32)
        cls = [np.sort((2 * rng.choice(act, 3, replace=False) + rng.integers(0, 2, 3)).astype(np.uint32)) for _ in range(12)]
        xl = np.concatenate(cls); xo = np.arange(0, 3 * 12 + 1, 3, dtype=np.uint64)
        xm = np.full(12, 1 | (1 << 4) | (3 << 6), np.uint32)
        extra = (xl, xo, xm)
    l2, o2, m2, vs2 = second_call_inputs(V, od1, extra)
    over = helpers.opts_from_flags(flags); over["sigma_calls"] = 2
    od2, _, _ = helpers.run_oracle(V, l2, o2, meta=m2, vstate=vs2, **over)
    s = sigma().Simplifier(0, flags=flags)
    try:
        s.load(V, lits, offs)
        fin1 = s.simplify()
        ed1 = to_dump(V, s.store(), fin1["cnfstate"])
        assert not sgd.compare(ed1, od1)
        view = s.device_view()
        assert view.max_var == V and view.live_clauses == fin1["clauses"] and view.headers and view.literals and view.ot_entries
        if extra is None:
            s.continue_resident()
        else:
            s.continue_resident(new_lits=extra[0], new_offs=extra[1], new_meta=extra[2])
        fin2 = s.simplify()
        ed2 = to_dump(V, s.store(), fin2["cnfstate"])
        assert not sgd.compare(ed2, od2), name
        assert (ed2.bits == od2.bits).all()
        assert s.memory()["cuda_mallocs"] == 1
    finally:
        s.close()


def test_compact_store_equals_full_store():
    fam, seed, args = SMALL["miter_x"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    s = sigma().Simplifier(0)
    try:
        s.load(V, lits, offs)
        fin = s.simplify()
        full = s.store()
        comp = s.store_compact()
        assert (comp["bits"] == full["bits"]).all() and (comp["lits"] == full["lits"]).all()
        assert (comp["sizes"] == np.diff(full["offs"].astype(np.int64)).astype(np.uint32)).all()
        assert (comp["eliminated"] == full["eliminated"]).all() and (comp["resolved"] == full["resolved"]).all()
    finally:
        s.close()


def test_load_with_32_bit_offsets_equals_load():
    """sigma_load32 (4-byte clause offsets over PCIe, widened on the device) gives the result of sigma_load, learnt marks included."""
    for name in ("miter_x", "k3_r30"):
        fam, seed, args = SMALL[name]
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        meta = np.zeros(len(offs) - 1, np.uint32)
        meta[::7] = 1 | (3 << 6)   # some learnt clauses (word 0: CB_LEARNT, lbd 3 as the later-call tests build them)
        res = []
        for o in (np.asarray(offs, np.uint64), np.asarray(offs, np.uint64).astype(np.uint32)):
            s = sigma().Simplifier(0)
            try:
                s.load(V, lits, o, meta=meta if name == "k3_r30" else None)
                s.simplify()
                res.append(s.store())
            finally:
                s.close()
        for k in ("bits", "sig", "offs", "lits", "eliminated"):
            assert (res[0][k] == res[1][k]).all(), (name, k)
        assert sorted(res[0]["resolved"].tolist()) == sorted(res[1]["resolved"].tolist())


def test_trail_ranges_split_seed_units_from_derived_ones():
    """sigma_trail_info: per prop() the units SUB/BVE produced come first (enqueueDevUnit), the derived ones after
    (enqueueUnit), elimbcp.cu:185-200."""
    rng = np.random.default_rng(5)
    found = False
    for seed in range(40):
        V = 120
        lits, offs = random_cnf(np.random.default_rng(900 + seed), V, 420, 2, 4)
        s = sigma().Simplifier(0)
        try:
            s.load(V, lits, offs)
            s.begin()
            total = 0
            while True:
                rep, done = s.round()
                if rep["propagated"] and rep["trail_added"]:      # (a prop() that ends in a conflict leaves no trail range)
                    info = s.trail_info()
                    assert info["last_from"] == total and info["last_count"] == rep["trail_added"] >= info["last_seeds"] > 0
                    assert info["last_seeds"] == min(rep["propagated"], rep["trail_added"])
                    total += info["last_count"]
                    found = True
                if done:
                    break
            fin = s.finish()
            if fin["cnfstate"] == 2:
                assert s.trail_info()["total"] == total == fin["trail_units"]
        finally:
            s.close()
    assert found


def test_late_option_change_that_needs_more_room_is_refused_until_reload():
    fam, seed, args = SMALL["k3_r42"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    S = sigma()
    s = S.Simplifier(0, flags=["-no-ve", "-no-veextend"])       # no room for resolvents is carved
    try:
        s.load(V, lits, offs)
        s.simplify()
        s.optSimp(flags=[])                                      # BVE on: needs the larger arena
        with pytest.raises(S.SigmaError):
            s.simplify()
        s.load(V, lits, offs)
        od, _, _ = helpers.run_oracle(V, lits, offs)
        fin = s.simplify()
        assert not sgd.compare(to_dump(V, s.store(), fin["cnfstate"]), od)
    finally:
        s.close()


def test_bad_literal_is_rejected_not_dereferenced():
    lits = np.array([2, 4, 6, 2, 5, 4000], np.uint32)     # literal 4000 with max_var 3
    offs = np.array([0, 3, 6], np.uint64)
    S = sigma()
    s = S.Simplifier(0)
    try:
        s.load(3, lits, offs)
        with pytest.raises(S.SigmaError):
            s.simplify()
    finally:
        s.close()


def test_reduction_log_tables_match_the_round_reports():
    """LOGREDALL / LOGREDCL (logging.hpp:152-158): per-stage survivors; the last stage of a round equals the round report,
    BVE's removed variables equal the round's eliminated count, and switching the log on changes nothing else."""
    fam, seed, args = SMALL["miter_a"]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    S = sigma()
    s = S.Simplifier(0, flags=["-all", "--verbose=2"])
    s0 = S.Simplifier(0, flags=["-all"])
    try:
        s.load(V, lits, offs); s0.load(V, lits, offs)
        fin, fin0 = s.simplify(), s0.simplify()
        assert (fin["clauses"], fin["literals"], fin["eliminated_vars"]) == (fin0["clauses"], fin0["literals"], fin0["eliminated_vars"])
        assert s0.reduction_log() == []
        log, rounds = s.reduction_log(), s.rounds()
        assert {e["stage"] for e in log} >= {"SUB", "BVE", "BCE", "ERE"}
        for r in rounds:
            mine = [e for e in log if e["round"] == r["round"] and e["stage"] != "BCP"]
            if r["kind"] == 0:
                assert [e["stage"] for e in mine] == ["SUB", "BVE", "BCE"][: len(mine)]
                assert mine[-1]["clauses"] == r["clauses"] and mine[-1]["literals"] == r["literals"]
                assert [e for e in mine if e["stage"] == "BVE"][0]["vars_removed"] == r["eliminated"]
                assert all(a["clauses"] >= b["clauses"] or b["stage"] == "BVE" for a, b in zip(mine, mine[1:]))
            elif r["kind"] == 1 and mine:
                assert mine[-1]["stage"] == "ERE" and mine[-1]["clauses"] == r["clauses"]
    finally:
        s.close(); s0.close()
