#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference GPU solver (runs on the B200 box).

    /usr/local/graft/bin/gpurun --timeout 1500 -- 'python tests/golden/make_golden.py'

For every (instance, flag-variant) below it
  1. regenerates the CNF with tools/cnfgen.cpp (seeded, deterministic),
  2. runs oracle/_ref/ref_driver (reference objects + our dump main, see oracle/ref/) on it,
  3. stores the SGD1 dump (small instances, gzip) and its fingerprint (all instances)
under gpurun_out/golden/.  The dumps + summary.json are then committed under tests/golden/
and pin oracle/sigma_oracle.cpp and the CUDA engine (tests/test_golden.py).

`--phases=K -no-ere` gives "the state after K SUB/BVE rounds" (the reference has no per-round
dump hook); the default variant adds the final ERE-only round.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sgd  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
TMP = "/tmp/golden_cnf"

# name -> (family, seed, args)
SMALL = {
    "k3_r42": ("ksat", 11, [600, 2520, 3]),
    "k3_r30": ("ksat", 12, [800, 2400, 3]),
    "k5_r10": ("ksat", 13, [400, 4000, 5]),
    "k4_r7": ("ksat", 14, [500, 3500, 4]),
    "miter_x": ("miter", 21, [30, 600, 900, 100, 8]),
    "miter_a": ("miter", 22, [40, 700, 300, 200, 8]),
    "mult6": ("mult", 31, [6]),
    "mult10": ("mult", 32, [10]),
    "parity": ("parity", 41, [300]),
    "multpar": ("multpar", 51, [5, 120]),
}
MEDIUM = {  # fingerprints only
    "cfg1_k3_100k": ("ksat", 1, [100000, 426000, 3]),
    "miter_50k": ("miter", 3, [2000, 50000, 900, 100, 32]),
    "mult48": ("mult", 4, [48]),
    "k5_20k": ("ksat", 2, [20000, 200000, 5]),
}
BASE = ["-no-lcvefast", "-quiet"]
VARIANTS = {
    "p1": ["--phases=1", "-no-ere"],
    "p2": ["--phases=2", "-no-ere"],
    "p3": ["--phases=3", "-no-ere"],
    "p4": ["--phases=4", "-no-ere"],
    "p5": ["--phases=5", "-no-ere"],
    "def": [],
    "nofun": ["-no-vefunction"],
    "nofun_p2": ["-no-vefunction", "--phases=2", "-no-ere"],
    "bce": ["-bce"],
    "all": ["-all"],
    "p1_ere": ["--phases=1"],
    "nosub_p2": ["-no-sub", "-no-veextend", "--phases=2", "-no-ere"],
    "aggr_p3": ["-aggresivesort", "--phases=3", "-no-ere"],
    "aggr_bce_p2": ["-aggresivesort", "-bce", "--phases=2", "-no-ere"],
}
# -lcvefast (the reference CLI's default election): the elected SET is deterministic, the order of elected[] and of the
# frozen list comes from atomics (lcve.cu:204-217) - pinned as multisets, with -no-vefunction (function-table indices
# follow the frozen-list order).  `--fast` runs only these.
FAST_VARIANTS = {
    "fast_nofun_p1": ["-lcvefast", "-no-vefunction", "--phases=1", "-no-ere"],
    "fast_nofun_p3": ["-lcvefast", "-no-vefunction", "--phases=3", "-no-ere"],
    "fast_nofun": ["-lcvefast", "-no-vefunction"],
}
MEDIUM_VARIANTS = ["p1", "p2", "def", "nofun"]
# The reference's ERE launch is broken whenever numElected < 4 * 8 * #SMs (every small instance): ereAsync sizes the
# dynamic shared memory for block.y = ereminthreads = 4 rows (elimination.cu:206), then LOOP_OVERSUB_Y raises block.y
# to max(ereminthreads, warpSize) = 32 (grid.cuh:85-104 with minThreads = MAX(MINTHREADS, warp), grid.cuh:196) without
# resizing it, and ere_k's rows 4..31 write their resolvents past the allocation (redundancy.cuh:160) -> illegal
# shared-memory access, sticky context error, garbage dump.  With --ereminthreads=32 (an option of the unmodified
# binary) the size is computed for 32 rows from the start and the tuner leaves the shape alone.  Confirmed with
# compute-sanitizer on the B200 (profiles/r02_ref_ere_sanitizer.log).  The launch shape does not enter ERE's result.
ERE_LAUNCH_FIX = ["--ereminthreads=32"]


def sh(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)


def main():
    """`--only v1,v2` runs just those variants (small instances) and merges them into the summary."""
    only = None
    ere_only = "--ere-only" in sys.argv     # re-run just the variants whose last round is ERE (with ERE_LAUNCH_FIX)
    for i, a in enumerate(sys.argv):
        if a == "--only":
            only = sys.argv[i + 1].split(",")
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(TMP, exist_ok=True)
    gen = os.path.join(ROOT, "build", "cnfgen")
    os.makedirs(os.path.dirname(gen), exist_ok=True)
    r = sh(["g++", "-O2", "-DCNFGEN_MAIN", "-o", gen, os.path.join(ROOT, "tools", "cnfgen.cpp")])
    assert r.returncode == 0, r.stdout
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    summary = {}
    if only:
        summary = json.load(open(os.path.join(ROOT, "tests", "golden", "summary.json")))
    log = open(os.path.join(OUT, "runs.log"), "w")
    groups = ((SMALL, only, True),) if only else ((SMALL, list(VARIANTS), True), (MEDIUM, MEDIUM_VARIANTS, False))
    if "--fast" in sys.argv:
        summary = json.load(open(os.path.join(ROOT, "tests", "golden", "summary.json")))
        VARIANTS.update(FAST_VARIANTS)
        groups = ((SMALL, list(FAST_VARIANTS), False), (MEDIUM, ["fast_nofun_p1", "fast_nofun"], False))
    if ere_only:
        summary = json.load(open(os.path.join(ROOT, "tests", "golden", "summary.json")))
        ere = [v for v, f in VARIANTS.items() if "-no-ere" not in f]
        groups = ((SMALL, ere, True), (MEDIUM, [v for v in MEDIUM_VARIANTS if v in ere], False))
    for group, variants, keep in groups:
        for name, (fam, seed, args) in group.items():
            cnf = os.path.join(TMP, name + ".cnf")
            r = sh([gen, fam, str(seed), cnf] + [str(a) for a in args])
            assert r.returncode == 0, r.stdout
            for var in variants:
                key = f"{name}__{var}"
                dump = os.path.join(TMP, key + ".sgd")
                if os.path.exists(dump):
                    os.remove(dump)
                flags = BASE + VARIANTS[var]
                if "-lcvefast" in flags:
                    flags = [f for f in flags if f != "-no-lcvefast"]
                if "-no-ere" not in flags:
                    flags = flags + ERE_LAUNCH_FIX
                t0 = time.time()
                try:
                    r = sh([drv, cnf, dump] + flags, timeout=600)
                    rc, out = r.returncode, r.stdout
                except subprocess.TimeoutExpired as e:
                    rc, out = -9, (e.stdout or "") + "\nTIMEOUT"
                dt = time.time() - t0
                log.write(f"=== {key} rc={rc} {dt:.2f}s flags={' '.join(flags)}\n{out[-3000:]}\n")
                log.flush()
                entry = {"family": fam, "seed": seed, "args": args, "flags": flags, "rc": rc, "wall_s": round(dt, 3)}
                for line in out.splitlines():
                    if line.startswith("s "):
                        entry["answer"] = line[2:].strip()
                    if "ref_driver: simplify wall" in line:
                        entry["simplify_ms"] = float(line.split("wall")[1].split("ms")[0])
                if os.path.exists(dump):
                    d = sgd.Dump.load(dump)
                    entry["fingerprint"] = d.fingerprint()
                    if keep:
                        d.save(os.path.join(OUT, key + ".sgd.gz"))
                summary[key] = entry
                print(key, rc, f"{dt:.2f}s", entry.get("fingerprint", {}).get("clauses"), flush=True)
    with open(os.path.join(OUT, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    log.close()


if __name__ == "__main__":
    main()
