#!/usr/bin/env python
"""Fingerprints of the CPU oracle (oracle/sigma_oracle.cpp, itself pinned on the reference's dumps)
on the BASELINE.json configs at FULL size, computed in the build container (minutes of CPU time) and
committed as tests/golden/fullsize_fingerprints.json, so that the GPU box only has to run the engine
and compare (tests/test_gpu_parity.py::test_full_size_matches_oracle_fingerprint).

    python tests/golden/make_fullsize_fingerprints.py [cfg1 cfg4 cfg2 cfg3]
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402

OUT = os.path.join(HERE, "fullsize_fingerprints.json")
KEYS = ["cnfstate", "clauses", "literals", "eliminated", "forced", "resolved_words", "resolved_groups", "trail",
        "h_lits_multiset", "h_full_multiset", "h_lits_ordered", "h_full_ordered", "h_eliminated", "h_forced",
        "h_resolved_groups", "h_resolved_records", "h_trail_multiset"]


def main():
    """argv: config names, optionally `name:flag,flag` for a flag variant (key `name__flag_flag`)."""
    specs = [a for a in sys.argv[1:] if a.split(":")[0] in helpers.CONFIGS] or ["cfg1", "cfg4", "cfg2", "cfg3"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for spec in specs:
        cfg, _, fl = spec.partition(":")
        flags = [f for f in fl.split(",") if f]
        name = cfg if not flags else cfg + "__" + "_".join(f.strip("-").replace("=", "") for f in flags)
        fam, seed, args = helpers.CONFIGS[cfg]
        t0 = time.time()
        V, lits, offs = helpers.gen_cnf(fam, seed, args)
        od, stats, _ = helpers.run_oracle(V, lits, offs, **helpers.opts_from_flags(flags))
        fp = od.fingerprint()
        res[name] = {"family": fam, "seed": seed, "args": list(args), "vars": int(V), "clauses_in": len(offs) - 1, "literals_in": len(lits),
                     "flags": flags, "rounds": [[int(x) for x in r] for r in stats],
                     "fingerprint": {k: fp[k] for k in KEYS}, "oracle_seconds": round(time.time() - t0, 1)}
        print(name, res[name]["oracle_seconds"], "s", res[name]["fingerprint"]["clauses"], "clauses", flush=True)
        json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
