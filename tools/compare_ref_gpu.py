#!/usr/bin/env python
"""Side-by-side timing on the B200 box: this engine vs the UNMODIFIED reference GPU build
(oracle/_ref/ref_driver = reference objects + a dump main, oracle/ref/) on the BASELINE configs.

    python tools/compare_ref_gpu.py [cfg1 cfg2 cfg3 cfg4] [--no-ref] [--fixed] [--out gpurun_out/compare.json]

Per config: generate the CNF (seeded), run the engine (device-resident + end-to-end, per-kernel
CUDA-event times, per-round reports), then the reference with `-profilegpu` in its default mode
(-lcvefast, the fair speed row) and - with --fixed - in the fixed-order mode the parity runs use
(-no-lcvefast, single-thread election).  Result fingerprints are compared when the modes match.
"""
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import cnfgen  # noqa: E402

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def run_engine(V, lits, offs, flags, steps=3):
    from parafrost_b200 import sigma
    s = sigma.Simplifier(0, flags=flags)
    t0 = time.perf_counter()
    s.load(V, lits, offs)
    t_load = time.perf_counter() - t0
    s.simplify()  # warm-up
    s.kernel_profile(1)
    reps = [s.simplify() for _ in range(steps)]
    kt = s.kernel_times()
    s.kernel_profile(0)
    rounds = s.rounds()
    t0 = time.perf_counter()
    s.load(V, lits, offs)
    rep = s.simplify()
    st = s.store()
    t_e2e = time.perf_counter() - t0
    out = {
        "ms_device": float(np.mean([r["ms_device"] for r in reps])), "ms_host": float(np.mean([r["ms_total"] for r in reps])),
        "ms_e2e_pageable": t_e2e * 1e3, "ms_load_pageable": t_load * 1e3, "launches": reps[-1]["kernel_launches"],
        "rounds": rounds, "clauses": rep["clauses"], "literals": rep["literals"], "eliminated": rep["eliminated_vars"], "cnfstate": rep["cnfstate"],
        "kernels": {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])},
        "memory": s.memory(),
    }
    s.close()
    return out, st


def run_ref(cnf, flags, timeout):
    dump = cnf + ".sgd"
    t0 = time.perf_counter()
    try:
        r = subprocess.run([DRV, cnf, dump, "-quiet", "-profilegpu"] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
        out, rc = r.stdout, r.returncode
    except subprocess.TimeoutExpired as e:
        out, rc = (e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\nTIMEOUT", -9
    wall = time.perf_counter() - t0
    res = {"rc": rc, "wall_s": wall, "flags": flags}
    m = re.search(r"simplify wall ([0-9.]+) ms, state (\d+), clauses (\d+)", out)
    if m:
        res.update({"simplify_ms": float(m.group(1)), "cnfstate": int(m.group(2)), "clauses": int(m.group(3))})
    m = re.search(r"stage ms (.*)", out)
    if m:
        tok = m.group(1).split()
        res["stage_ms"] = {tok[i]: float(tok[i + 1]) for i in range(0, len(tok) - 1, 2)}
    if "CUDA ERROR" in out:
        res["cuda_error"] = [l for l in out.splitlines() if "CUDA ERROR" in l][0][:200]
    if rc != 0 and "simplify_ms" not in res:
        res["tail"] = out[-600:]
    fp = None
    if os.path.exists(dump):
        try:
            import sgd
            if os.path.getsize(dump) < 1.5e9:
                fp = sgd.Dump.load(dump).fingerprint()
        except Exception as e:  # noqa: BLE001
            res["dump_error"] = repr(e)[:200]
        os.remove(dump)
    res["fingerprint"] = fp
    return res


def main():
    names = [a for a in sys.argv[1:] if a in cnfgen.CONFIGS] or ["cfg1", "cfg4", "cfg2", "cfg3"]
    no_ref = "--no-ref" in sys.argv
    fixed = "--fixed" in sys.argv
    outp = os.path.join(ROOT, "gpurun_out", "compare.json")
    for i, a in enumerate(sys.argv):
        if a == "--out":
            outp = sys.argv[i + 1]
    os.makedirs(os.path.dirname(outp), exist_ok=True)
    result = {}
    for name in names:
        fam, seed, gargs = cnfgen.CONFIGS[name]
        cnf = f"/tmp/{name}.cnf"
        t0 = time.perf_counter()
        V, lits, offs = cnfgen.gen_cnf(fam, seed, gargs, dimacs_path=None if no_ref else cnf)
        entry = {"family": fam, "seed": seed, "args": gargs, "vars": V, "clauses": len(offs) - 1, "literals": len(lits), "gen_s": time.perf_counter() - t0}
        eng, st = run_engine(V, lits, offs, [])
        entry["engine"] = eng
        if not no_ref:
            # the ERE kernel of the reference faults on sm_100: time it with and without
            entry["ref_default"] = run_ref(cnf, [], 900)
            entry["ref_noere"] = run_ref(cnf, ["-no-ere"], 900)
            eng2, st2 = run_engine(V, lits, offs, ["-no-ere"], steps=2)
            entry["engine_noere"] = eng2
            if fixed:
                entry["ref_fixed_noere"] = run_ref(cnf, ["-no-lcvefast", "-no-ere"], 1500)
                fp = entry["ref_fixed_noere"].get("fingerprint")
                if fp:
                    import sgd
                    ed = sgd.Dump.from_arrays(V, eng2["cnfstate"], st2["bits"], st2["sig"], st2["offs"], st2["lits"], st2["eliminated"], st2["resolved"], st2["trail"])
                    efp = ed.fingerprint()
                    keys = ["clauses", "literals", "eliminated", "h_lits_multiset", "h_full_ordered", "h_eliminated", "h_resolved_groups", "h_trail_multiset"]
                    entry["parity_vs_ref_fixed_noere"] = {k: efp[k] == fp[k] for k in keys}
            os.remove(cnf)
        result[name] = entry
        with open(outp, "w") as f:
            json.dump(result, f, indent=1)
        e = entry["engine"]
        print(f"{name}: engine {e['ms_device']:.2f} ms device ({len(e['rounds'])} rounds, {e['launches']} launches), "
              f"ref default {entry.get('ref_default', {}).get('simplify_ms')} ms, ref -no-ere {entry.get('ref_noere', {}).get('simplify_ms')} ms, "
              f"engine -no-ere {entry.get('engine_noere', {}).get('ms_device')}", flush=True)
        top = list(e["kernels"].items())[:6]
        print("   top kernels: " + ", ".join(f"{k} {v['ms_per_step']:.2f}" for k, v in top), flush=True)


if __name__ == "__main__":
    main()
