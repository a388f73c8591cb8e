#!/usr/bin/env python
"""One simplification of a BASELINE config, nothing else - the target command for ncu captures.
    ncu ... python tools/profile_run.py cfg2 [reference CLI flags]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cnfgen  # noqa: E402
from parafrost_b200 import sigma  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
fam, seed, args = cnfgen.CONFIGS[name]
V, lits, offs = cnfgen.gen_cnf(fam, seed, args)
s = sigma.Simplifier(0, flags=sys.argv[2:])
s.load(V, lits, offs)
rep = s.simplify()
print(name, {k: rep[k] for k in ("rounds", "clauses", "literals", "eliminated_vars", "kernel_launches", "ms_device")})
for r in s.rounds():
    print("  ", r)
s.close()
