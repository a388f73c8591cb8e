#!/usr/bin/env python
"""A/B timing of engine variants on one workload (GPU box): device ms per simplify() and the top kernels.
    SIGMA_VE_LOCAL=0 python tools/kernel_ab.py cfg3 [steps] [--check]     # --check: result fingerprint, to compare variants
Variants are chosen by environment variables read by the engine (csrc/elim.cu, cnf.cu)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402

import cnfgen  # noqa: E402
from parafrost_b200 import sigma  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
    flags = [a for a in sys.argv[2:] if a.startswith("-") and a != "--check"]
    fam, seed, args = cnfgen.CONFIGS[name]
    V, lits, offs = cnfgen.gen_cnf(fam, seed, args)
    s = sigma.Simplifier(0, flags=flags)
    s.load(V, lits, offs)
    s.simplify()
    ms = [s.simplify()["ms_device"] for _ in range(steps)]
    s.kernel_profile(1)
    rep = s.simplify()
    ks = s.kernel_stats()
    s.kernel_profile(0)
    out = {"workload": name, "env": {k: v for k, v in os.environ.items() if k.startswith("SIGMA_")}, "ms_device": float(np.mean(ms)),
           "ms_min": float(np.min(ms)), "launches": rep["kernel_launches"], "rounds": [(r["kind"], r["elected"], r["eliminated"]) for r in s.rounds()],
           "clauses": rep["clauses"], "literals": rep["literals"], "eliminated": rep["eliminated_vars"],
           "top": [(k, round(v[0], 3), v[1], round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[2] and v[0] else None) for k, v in sorted(ks.items(), key=lambda kv: -kv[1][0])[:14]]}
    if "--check" in sys.argv:
        st = s.store()
        h = hashlib.md5()
        for k in ("bits", "sig", "offs", "lits", "eliminated"):
            h.update(np.ascontiguousarray(st[k]).tobytes())
        out["md5_ordered"] = h.hexdigest()
        out["resolved_words"] = int(len(st["resolved"]))
    s.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
