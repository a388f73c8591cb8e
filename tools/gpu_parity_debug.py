#!/usr/bin/env python
"""Round-by-round comparison of the CUDA engine with the CPU oracle (runs on the GPU box).

    python tools/gpu_parity_debug.py [name ...]

Prints, per instance and flag variant, the first round where the engine's live clause list,
elected set or counters differ from the oracle's, with enough detail to locate the kernel.
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
import sgd  # noqa: E402
from parafrost_b200 import sigma  # noqa: E402

INSTANCES = {
    "k3_r30": ("ksat", 12, [800, 2400, 3]),
    "k3_r42": ("ksat", 11, [600, 2520, 3]),
    "k5_r10": ("ksat", 13, [400, 4000, 5]),
    "miter_x": ("miter", 21, [30, 600, 900, 100, 8]),
    "miter_a": ("miter", 22, [40, 700, 300, 200, 8]),
    "mult6": ("mult", 31, [6]),
    "mult10": ("mult", 32, [10]),
    "parity": ("parity", 41, [300]),
    "multpar": ("multpar", 51, [5, 120]),
    "cfg1_k3_100k": ("ksat", 1, [100000, 426000, 3]),
    "miter_50k": ("miter", 3, [2000, 50000, 900, 100, 32]),
    "mult48": ("mult", 4, [48]),
}
VARIANTS = {
    "p5": ["--phases=5", "-no-ere"],
    "def": [],
    "nofun": ["-no-vefunction"],
    "all": ["-all"],
    "nosub_p2": ["-no-sub", "-no-veextend", "--phases=2", "-no-ere"],
}


def to_dump(V, st, state=2):
    return sgd.Dump.from_arrays(V, state, st["bits"], st["sig"], st["offs"], st["lits"], st["eliminated"], st["resolved"], st["trail"])


def describe_diff(a, b, limit=5):
    """a = engine dump, b = oracle dump"""
    out = []
    ca, cb = a.ordered_clauses(), b.ordered_clauses()
    out.append(f"    clauses engine {len(ca)} oracle {len(cb)}")
    sa, sb = {}, {}
    for c in ca:
        sa[c] = sa.get(c, 0) + 1
    for c in cb:
        sb[c] = sb.get(c, 0) + 1
    only_a = [c for c in sa if sa[c] > sb.get(c, 0)]
    only_b = [c for c in sb if sb[c] > sa.get(c, 0)]
    out.append(f"    only in engine: {len(only_a)}  e.g. {only_a[:limit]}")
    out.append(f"    only in oracle: {len(only_b)}  e.g. {only_b[:limit]}")
    if not only_a and not only_b:
        for i, (x, y) in enumerate(zip(ca, cb)):
            if x != y:
                out.append(f"    same multiset, first order difference at {i}: engine {x} oracle {y}")
                break
        nb = np.nonzero(a.bits != b.bits)[0]
        ns = np.nonzero(a.sig != b.sig)[0]
        if len(nb):
            i = int(nb[0])
            out.append(f"    bits differ at {len(nb)} clauses, first {i}: engine {int(a.bits[i]):#x} oracle {int(b.bits[i]):#x} clause {ca[i]}")
        if len(ns):
            i = int(ns[0])
            out.append(f"    sig differ at {len(ns)} clauses, first {i}: engine {int(a.sig[i]):#x} oracle {int(b.sig[i]):#x} clause {ca[i]}")
    return "\n".join(out)


def run_one(name, var, verbose=True):
    fam, seed, args = INSTANCES[name]
    flags = VARIANTS[var]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, osnaps = helpers.run_oracle(V, lits, offs, snapshots=True, **helpers.opts_from_flags(flags))
    elections = od.extra.get("elections", [])
    s = sigma.Simplifier(0, flags=flags)
    s.load(V, lits, offs)
    t0 = time.time()
    s.begin()
    ok = True
    r = 0
    while True:
        rep, done = s.round()
        if rep["kind"] != 2 or rep["elected"]:
            # election check (survivors after BVE when kind == 0, so compare counts only there)
            if r < len(elections) and rep["elected"] != len(elections[r]):
                print(f"  [{name}/{var}] round {r}: elected {rep['elected']} != oracle {len(elections[r])}")
                ok = False
        if rep["kind"] == 0:
            if r < len(ors):
                exp = ors[r]
                got = (rep["elected"], rep["eliminated"], rep["resolvents"], rep["clauses"], rep["literals"])
                if tuple(int(x) for x in exp) != got:
                    print(f"  [{name}/{var}] round {r}: (elected, eliminated, resolvents, clauses, literals) engine {got} oracle {tuple(int(x) for x in exp)}")
                    ok = False
                snap = to_dump(V, s.snapshot())
                d = sgd.compare(snap, osnaps[r]) if r < len(osnaps) else ["no oracle snapshot"]
                d = [x for x in d if not x.startswith(("eliminated", "forced", "resolved", "trail", "h_elim", "h_forced", "h_resolved", "h_trail"))]
                if d:
                    print(f"  [{name}/{var}] round {r}: snapshot differs: {d[:4]}")
                    print(describe_diff(snap, osnaps[r]))
                    ok = False
            else:
                print(f"  [{name}/{var}] round {r}: engine ran an extra round {rep}")
                ok = False
            r += 1
        if not ok or done:
            break
    fin = s.finish()
    dt = time.time() - t0
    if ok:
        if r != len(ors):
            print(f"  [{name}/{var}] engine ran {r} rounds, oracle {len(ors)}")
            ok = False
        ed = to_dump(V, s.store(), fin["cnfstate"])
        d = sgd.compare(ed, od)
        if d:
            print(f"  [{name}/{var}] FINAL differs: {d[:6]}")
            if any(x.startswith(("clauses", "literals", "h_lits", "h_full")) for x in d):
                print(describe_diff(ed, od))
            if any("resolved" in x for x in d):
                ga, gb = set(ed.resolved_groups()), set(od.resolved_groups())
                print(f"    resolved groups only engine {list(ga - gb)[:3]} only oracle {list(gb - ga)[:3]}")
            if any("trail" in x for x in d):
                print(f"    trail engine {sorted(ed.trail.tolist())[:10]} oracle {sorted(od.trail.tolist())[:10]}")
            if any("elim" in x for x in d):
                ea, eb = set(ed.eliminated_vars()), set(od.eliminated_vars())
                print(f"    eliminated only engine {sorted(ea - eb)[:10]} only oracle {sorted(eb - ea)[:10]}")
            ok = False
    print(f"{'OK  ' if ok else 'FAIL'} {name}/{var}: rounds {r}, clauses {fin['clauses']}, state {fin['cnfstate']}, "
          f"launches {fin['kernel_launches']}, {dt * 1e3:.1f} ms")
    s.close()
    return ok


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(INSTANCES)
    variants = list(VARIANTS)
    nfail = 0
    for name in names:
        for var in variants:
            if name in ("cfg1_k3_100k", "miter_50k", "mult48") and var not in ("p5", "def"):
                continue
            try:
                nfail += not run_one(name, var)
            except Exception as e:  # keep going: one call = many diagnostics
                nfail += 1
                print(f"EXC  {name}/{var}: {e!r}")
                traceback.print_exc()
            sys.stdout.flush()
    print(f"failures: {nfail}")


if __name__ == "__main__":
    main()
