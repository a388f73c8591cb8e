#!/usr/bin/env python
"""Debug helper (GPU box): engine vs oracle witness groups on one small case with learnt clauses."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers, sgd
from parafrost_b200 import sigma as S

name, flags, calls = sys.argv[1], [f for f in sys.argv[2].split(",") if f], int(sys.argv[3])
SMALL = {"mult10": ("mult", 32, [10]), "k3_r30": ("ksat", 12, [800, 2400, 3]), "miter_a": ("miter", 22, [40, 700, 300, 200, 8]), "k4_r7": ("ksat", 14, [500, 3500, 4])}
rng = np.random.default_rng(23)
for nm in ("k3_r30", "miter_a", "mult10", "k4_r7"):   # same rng stream as the test
    fam, seed, args = SMALL[nm]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    meta = np.zeros(len(offs) - 1, np.uint32)
    lrn = rng.random(len(meta)) < 0.2
    meta[lrn] = 1 | (rng.integers(0, 3, int(lrn.sum())).astype(np.uint32) << 4) | (rng.integers(2, 9, int(lrn.sum())).astype(np.uint32) << 6)
    if nm == name:
        break
over = helpers.opts_from_flags(flags); over["sigma_calls"] = calls
od, ors, _ = helpers.run_oracle(V, lits, offs, meta=meta, **over)
s = S.Simplifier(0, flags=flags, sigma_calls=calls)
s.load(V, lits, offs, meta=meta)
fin = s.simplify()
st = s.store()
ed = sgd.Dump.from_arrays(V, fin["cnfstate"], st["bits"], st["sig"], st["offs"], st["lits"], st["eliminated"], st["resolved"], st["trail"])
print("rounds engine", [(r["kind"], r["elected"], r["eliminated"], r["resolvents"], r["clauses"], r["literals"]) for r in s.rounds()])
print("rounds oracle", ors.tolist())
ge, go = ed.resolved_groups(), od.resolved_groups()
print("groups", len(ge), len(go), "words", len(ed.resolved), len(od.resolved))
se, so = set(ge), set(go)
print("only engine:", sorted(se - so)[:10])
print("only oracle:", sorted(so - se)[:10])
clauses = {}
for i in range(len(offs) - 1):
    c = tuple(lits[int(offs[i]):int(offs[i + 1])].tolist())
    clauses.setdefault(frozenset(c), []).append((i, int(meta[i])))
for g in sorted(se ^ so)[:6]:
    print("group", g)
