#!/usr/bin/env python
"""bench.py -- the SIGmA inprocessing hot path on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A *step* is one complete simplification (Solver::simplify(): awaken + every round of
histogram/scan/scatter -> election -> list sort -> SUB/BVE/(BCE)/ERE -> GC) of one synthetic CNF
of the named BASELINE.json shape.  Metric: literals processed per second
  = sum over rounds of the live literals at the start of the round  /  time of the step
(SURVEY.md 8d), reported with wall-ms per round beside it.

  value  inputs already resident in HBM (sigma_load done); K x sigma_run timed with CUDA events
         on the engine's launch stream.
  e2e    the reference-facing call sequence with HOST buffers: sigma_load (pinned host -> HBM),
         sigma_run, sigma_store (HBM -> pinned host), all inside the timed region.
  roofline      dominant kernel of the timed region: algorithmic bytes / CUDA-event duration
                (event pair around every launch on the launch stream) vs MEASURED_PEAKS.json.
  cpu_baseline  the UNMODIFIED reference CPU simplifier (oracle/_ref/parafrost_cpu, built from
                /root/reference by oracle/ref/Makefile), single-threaded, on a bounded sample.

Multi-GPU: the path does not shard - N ranks simplify N independent CNFs of the
same shape (seed + rank): weak scaling, no collective on the data path; torch.distributed (NCCL)
only provides the barrier and the max-over-ranks reduction of the timings.

--impl reference times the reference's own CPU implementation on the host cores (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

METRIC = "simplify_literals_per_s"
UNIT = "literals/s"
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "parafrost_cpu")


# ----------------------------------------------------------------------------- workloads
def workload_spec(name, rank=0, scale=1.0):
    import cnfgen
    fam, seed, args = cnfgen.CONFIGS[name]
    args = list(args)
    if scale != 1.0:
        if fam == "ksat":
            args[0] = max(10, int(args[0] * scale)); args[1] = max(10, int(args[1] * scale))
        elif fam == "miter":
            args[0] = max(4, int(args[0] * scale)); args[1] = max(8, int(args[1] * scale))
        elif fam == "multpar":
            args[0] = max(4, int(args[0] * scale ** 0.5)); args[1] = max(8, int(args[1] * scale))
    from parafrost_b200 import replicas
    return fam, replicas.rank_seed(seed, rank), args


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        """indices: the GPUs of the job.  ONE sampler (rank 0) watches all of them: a poller per rank means 8 nvidia-smi
        processes taking the driver's locks every 100 ms inside a 20 ms timed region."""
        self.index = ",".join(str(i) for i in (indices if isinstance(indices, (list, tuple, range)) else [indices]))
        self.p, self.lines = None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", self.index],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.p.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- roofline
# Algorithmic bytes per kernel come from the engine itself (sigma_kernel_stats, csrc/*.cu: KB(...) beside every launch;
# formulas in DESIGN.md 3 / SURVEY.md 8d) - the bench no longer keeps a second copy of them.
def ncu_traffic(workload):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of this workload's kernels
    from the newest committed `ncu --set full` capture (profiles/*_ncu_traffic_<workload>.json,
    written by tools/ncu_digest.py --json)."""
    import glob
    best = {}
    def version(f):   # rNN_..._vMM.json: by round, then by capture number (later captures override earlier ones)
        b = os.path.basename(f)
        r = re.match(r"r(\d+)_", b)
        m = re.search(r"_v(\d+)[a-z]*\.json$", b)
        return (int(r.group(1)) if r else -1, int(m.group(1)) if m else -1)
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_ncu_traffic_{workload}*.json")), key=version):
        try:
            best.update(json.load(open(f)).get("kernels", {}))
        except (OSError, ValueError):
            pass
    return best


def kbase(name):
    """'(k_ot_part<3, 5>)' / 'void k_sub<4>' -> 'k_ot_part' / 'k_sub' (the LAUNCH macro stringifies its argument)."""
    return name.replace("void ", "").strip("() ").split("<")[0]


# ----------------------------------------------------------------------------- reference CPU arm
def parse_ref_cpu(out, L0):
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    stage_ms, rounds_L, last_L = 0.0, [], L0
    in_report = False
    for ln in out.splitlines():
        if "Simplifier Report" in ln:
            in_report = True
            continue
        if in_report:
            m = re.match(r"c\s+-\s+(.+?)\s*:\s*([0-9.]+)\s+ms", ln)
            if m:
                stage_ms += float(m.group(2))
            elif "Sigmifications" in ln:
                in_report = False
        m = re.match(r"c\s+Survived\s*:\s*(\d+)\s+(\d+)\s+(\d+)", ln)
        if m:
            last_L = int(m.group(3))
        if "Electing variables in phase-" in ln:
            rounds_L.append(last_L)
    return stage_ms, rounds_L


def run_ref_cpu(cnf_path, L0, timeout=900):
    """One simplify() of the UNMODIFIED reference CPU solver (./install.sh -c equivalent build).
    Time = its own -profilesimp stage timers (simplifier only, parsing excluded)."""
    t0 = time.perf_counter()
    r = subprocess.run([REF_CPU, cnf_path, "-no-solve", "-profilesimp", "--verbose=2"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    ms, rounds_L = parse_ref_cpu(r.stdout, L0)
    if ms <= 0:
        raise RuntimeError("reference CPU run gave no simplifier report:\n" + r.stdout[-2000:])
    if not rounds_L:
        rounds_L = [L0]
    return {"ms": ms, "rounds": len(rounds_L), "literals": sum(rounds_L), "wall_s": wall}


def ref_sample(workload, budget_literals=20_000_000):
    """Bounded sample of the workload for the CPU legs: same family and clause/variable ratio,
    scaled so one simplify() is ~10-30 s of single-core work."""
    import cnfgen
    fam, seed, args = cnfgen.CONFIGS[workload]
    V, lits, offs = None, None, None
    # literals of the full config
    full_L = {"cfg1": 1_278_000, "cfg2": 105_000_000, "cfg3": 115_830_620, "cfg4": 26_904_824}[workload]
    scale = min(1.0, budget_literals / full_L)
    fam, seed, args = workload_spec(workload, 0, scale)
    path = f"/tmp/sigma_bench_{workload}_{os.getpid()}.cnf"
    V, lits, offs = cnfgen.gen_cnf(fam, seed, args, dimacs_path=path)
    desc = f"{fam}{tuple(args)} seed {seed}: V={V} C={len(offs) - 1} L={len(lits)} ({scale:.4g} of {workload})"
    return path, V, len(offs) - 1, len(lits), desc


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(REF_CPU):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/parafrost_cpu not built (make -f oracle/Makefile ref)"}))
        return 0
    path, V, Cn, L, desc = ref_sample(a.workload)
    try:
        for _ in range(a.warmup):
            run_ref_cpu(path, L)
        runs = [run_ref_cpu(path, L) for _ in range(a.steps)]
    finally:
        os.remove(path)
    ms = sum(r["ms"] for r in runs)
    lit = L * len(runs)            # literals of the input formula per simplify() call, as in the engine's arm
    value = lit / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "ms_per_round": ms / max(1, sum(r["rounds"] for r in runs)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "literals_per_round_s": sum(r["literals"] for r in runs) / (ms * 1e-3),
        "config": {"workload": a.workload, "sample": desc, "solver": "reference CPU v3.2.5 (src/cpu), -no-solve -profilesimp, default inprocessing",
                   "unit_definition": "literals of the input formula per simplify() call (all rounds) / time"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": desc,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- reference GPU arm
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
# largest member of each family the UNMODIFIED reference GPU build gets through on a B200 (it aborts above: DESIGN.md 4)
# above 699 050 clauses its live-clause count halves (count.cu:131-141 sizes a 1024-thread reduction that reduce.cuh:88-102 cannot
# finish, on GPUs with more than 128 SMs), the CNF is compacted to the wrong size and moderngpu's segmented sort reads out of bounds
REF_GPU_SAMPLE = {"cfg1": 1.0, "cfg2": 650_000 / 21_000_000, "cfg3": 650_000 / 39_601_471, "cfg4": 650_000 / 8_400_000}


def run_ref_gpu(cnf_path, flags, timeout=900, host_mode=True):
    """One simplify() of the unmodified reference GPU solver (oracle/_ref/ref_driver = its objects + a dump main).
    host_mode: with the write-back into the host clause database (newClause), i.e. its end-to-end call."""
    env = dict(os.environ)
    if host_mode:
        env["REF_DRIVER_HOST"] = "1"
    dump = cnf_path + ".sgd"
    t0 = time.perf_counter()
    r = subprocess.run([REF_GPU, cnf_path, dump, "-quiet", "-profilegpu", "--ereminthreads=32"] + flags, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env)
    wall = time.perf_counter() - t0
    if os.path.exists(dump):
        os.remove(dump)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
    res = {"rc": r.returncode, "wall_s": wall, "flags": flags}
    m = re.search(r"simplify wall ([0-9.]+) ms, state (\d+), clauses (\d+)", out)
    if m:
        res.update({"simplify_ms": float(m.group(1)), "cnfstate": int(m.group(2)), "clauses": int(m.group(3))})
    m = re.search(r"stage ms (.*)", out)
    if m:
        tok = m.group(1).split()
        res["stage_ms"] = {tok[i]: float(tok[i + 1]) for i in range(0, len(tok) - 1, 2)}
    if "simplify_ms" not in res:
        res["tail"] = out[-400:]
    return res


def ref_gpu_sample(workload):
    import cnfgen
    scale = REF_GPU_SAMPLE.get(workload)
    if scale is None:
        return None
    fam, seed, args = workload_spec(workload, 0, scale)
    path = f"/tmp/sigma_bench_refgpu_{workload}_{os.getpid()}.cnf"
    V, lits, offs = cnfgen.gen_cnf(fam, seed, args, dimacs_path=path)
    desc = f"{fam}{tuple(args)} seed {seed}: V={V} C={len(offs) - 1} L={len(lits)} ({scale:.4g} of {workload})"
    return path, V, lits, offs, desc


def ref_gpu_block(workload, local, fixed=False):
    """The reference's own GPU build beside the engine, end to end, on the largest member of the workload's family the
    reference gets through correctly on this GPU (north_star: "the reference's own GPU build on the same B200 is also shown")."""
    import numpy as _np
    from parafrost_b200 import sigma
    smp = ref_gpu_sample(workload)
    if smp is None:
        return {"unavailable": "no sample of this family"}
    path, V, lits, offs, desc = smp
    out = {"sample": desc, "why_not_full_size": "the reference GPU build halves its clause count above 699 050 clauses on GPUs with > 128 SMs and then "
                                                "crashes (DESIGN.md 4, profiles/r02_ref_gpu_k5_4M_*.log)"}
    try:
        for name, flags in (("engine_lcvefast", ["-lcvefast"]), ("engine_fixed_order", [])):
            s = sigma.Simplifier(local, flags=flags)
            best = None
            for _ in range(4):
                t0 = time.perf_counter()
                s.load(V, lits, offs)
                rep = s.simplify()
                st = s.store_compact()
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
            out[name] = {"ms_e2e": best, "ms_device": rep["ms_device"], "clauses_out": rep["clauses"], "eliminated_vars": rep["eliminated_vars"],
                         "note": "sigma_load (pageable numpy arrays) + sigma_run + sigma_store_compact, best of 4"}
            s.close()
        r = run_ref_gpu(path, [], timeout=600, host_mode=True)
        out["reference_default"] = {k: r.get(k) for k in ("simplify_ms", "clauses", "stage_ms", "rc", "tail") if r.get(k) is not None}
        out["reference_default"]["note"] = "its default mode (-lcvefast), simplify(false): extract + H2D + rounds + write-back into the host clause database"
        if "simplify_ms" in r:
            out["speedup_e2e_vs_reference_default"] = r["simplify_ms"] / out["engine_lcvefast"]["ms_e2e"]
            if r.get("stage_ms"):   # its own -profilegpu stage timers only (no arena creation, no host loops) against the engine's device time
                ssum = sum(r["stage_ms"].values())
                out["reference_default"]["stage_ms_sum"] = ssum
                out["speedup_device_vs_reference_stage_sum"] = ssum / out["engine_lcvefast"]["ms_device"]
        if fixed:
            r2 = run_ref_gpu(path, ["-no-lcvefast"], timeout=900, host_mode=True)
            out["reference_fixed_order"] = {k: r2.get(k) for k in ("simplify_ms", "clauses", "rc", "tail") if r2.get(k) is not None}
            if "simplify_ms" in r2:
                out["speedup_e2e_vs_reference_fixed_order"] = r2["simplify_ms"] / out["engine_fixed_order"]["ms_e2e"]
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:300]
    finally:
        os.remove(path)
    return out


def reference_gpu_arm(a):
    """--impl reference-gpu: the reference's own GPU build on the largest member of the workload's family it survives."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    smp = ref_gpu_sample(a.workload) if os.path.exists(REF_GPU) else None
    if smp is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/ref_driver not built, or no size of this family runs on the reference GPU build"}))
        return 0
    path, V, lits, offs, desc = smp
    try:
        runs = [run_ref_gpu(path, a.flags.split()) for _ in range(max(1, a.steps))]
    finally:
        os.remove(path)
    ok = [r for r in runs if "simplify_ms" in r]
    if not ok:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "reference GPU run failed: " + runs[-1].get("tail", "")[-200:]}))
        return 0
    ms = sum(r["simplify_ms"] for r in ok) / len(ok)
    value = len(lits) / (ms * 1e-3)
    print(json.dumps({"impl": "reference-gpu", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": len(ok), "warmup": 0, "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                      "config": {"workload": a.workload, "sample": desc, "solver": "reference GPU v4.1.1 (src/gpu), its default election mode (-lcvefast) unless --flags says otherwise, "
                                 "simplify(false): extract + H2D + rounds + write-back to the host clause database", "flags": a.flags.split()},
                      "stage_ms": ok[-1].get("stage_ms"), "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}}))
    return 0


# ----------------------------------------------------------------------------- config 5: instance-parallel batch
def batch_main(a, torch, dist, barrier, rank, world, local):
    """BASELINE.json config 5: a batch of mixed CNFs, one engine context per instance, instances
    dealt to the ranks longest-first (parafrost_b200/replicas.py), no collective on the data path.
    Metric CNFs/s; `value` counts the sigma_run time of each instance with its formula resident,
    `e2e` the whole load -> run -> store sequence from / to pinned host memory.  "scaling": "strong"
    (the batch is fixed, ranks share it)."""
    import cnfgen
    from parafrost_b200 import replicas, sigma
    specs = replicas.batch_specs(a.batch, a.scale)
    weights = [replicas.spec_weight(sp) for sp in specs]
    mine = replicas.assign_longest_first(weights, world)[rank]
    stream = torch.cuda.Stream()
    s = sigma.Simplifier(local)
    s.set_stream(stream.cuda_stream)
    inst = []
    for i in mine:   # generated once, kept in pinned host memory for every step
        fam, seed, args = specs[i]
        keep = []

        def alloc(n, dt, keep=keep):
            t = torch.empty(int(n), dtype={np.uint32: torch.int32, np.uint64: torch.int64}[dt], pin_memory=True)
            keep.append(t)
            return t.numpy().view(dt)
        V, lits, offs = cnfgen.gen_cnf(fam, seed, args, alloc=alloc)
        offs = offs32_of(offs, alloc)
        inst.append((V, lits, offs, keep))
    maxC = max((len(x[2]) - 1 for x in inst), default=1)
    maxL = max((len(x[1]) for x in inst), default=1)
    maxV = max((x[0] for x in inst), default=1)
    pin = []

    def palloc(n, dt):
        t = torch.empty(int(n), dtype={np.uint32: torch.int32, np.uint64: torch.int64}[dt], pin_memory=True)
        pin.append(t)
        return t.numpy().view(dt)

    def outbuf():
        return {"bits": palloc(2 * maxC + 16, np.uint32), "sizes": palloc(2 * maxC + 16, np.uint32), "lits": palloc(2 * maxL + 16, np.uint32),
                "eliminated": np.zeros(maxV + 1, np.uint8), "resolved": palloc(maxC + maxL + 2, np.uint32), "trail": palloc(3 * (maxV + 1), np.uint32)}
    out0 = outbuf()

    def one_pass():
        """-> (ms of sigma_run summed over my instances, ms of the whole pass, launches, literals, h2d, d2h)"""
        run_ms, launches, lit, h2d, d2h = 0.0, 0, 0, 0, 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for V, lits, offs, _ in inst:
            s.load(V, lits, offs)
            rep = s.simplify()
            st = s.store_compact(into=out0)
            run_ms += rep["ms_device"]; launches += rep["kernel_launches"]
            lit += len(lits)
            h2d += int(lits.nbytes + offs.nbytes); d2h += sum(int(v.nbytes) for v in st.values())
        e1.record(stream)
        stream.synchronize()
        return run_ms, e0.elapsed_time(e1), launches, lit, h2d, d2h

    for _ in range(a.warmup):
        one_pass()
    clocks = ClockSampler(range(world)) if rank == 0 else None   # one sampler for the job's GPUs
    barrier()
    if clocks:
        clocks.start()
    tot = [one_pass() for _ in range(a.steps)]
    barrier()
    clk = clocks.stop() if clocks else None
    # ---- the same passes through K contexts per GPU (replicas.Pipeline): copies of one instance overlap the kernels of
    # another; every instance is still loaded from and stored to pinned host memory
    piped_ms = None
    if a.pipeline > 1:
        bufs = [outbuf() for _ in range(a.pipeline)]
        jobs = [(V, lits, offs) for V, lits, offs, _ in inst]
        with replicas.Pipeline(local, depth=a.pipeline, compact=True) as pipe:
            pipe.run(jobs, lambda *_: None, bufs)          # warm-up
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                pipe.run(jobs, lambda *_: None, bufs)
            torch.cuda.synchronize()
            piped_ms = (time.perf_counter() - t0) * 1e3
        piped_ms, _ = replicas.reduce_timing(dist, piped_ms, 0.0, device="cuda")
    # ---- profiled pass over this rank's instances (outside the timed regions): kernel table with the engine's byte counts
    s.kernel_profile(1)
    for V, lits, offs, _ in inst:
        s.load(V, lits, offs)
        s.simplify()
    kstats = s.kernel_stats()
    s.kernel_profile(0)
    run_ms = sum(t[0] for t in tot); all_ms = sum(t[1] for t in tot)
    n_mine = len(inst) * a.steps
    run_ms, n_all = replicas.reduce_timing(dist, run_ms, float(n_mine), device="cuda")
    all_ms, lit_all = replicas.reduce_timing(dist, all_ms, float(sum(t[3] for t in tot)), device="cuda")
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        e2e_serial = n_all / (all_ms * 1e-3)
        e2e = {"value": e2e_serial, "unit": "CNFs/s", "ms_per_step": all_ms / a.steps, "mode": "one context per GPU", "h2d_bytes_per_step": tot[-1][4],
               "d2h_bytes_per_step": tot[-1][5]}
        if piped_ms:
            e2e["serial"] = {"value": e2e_serial, "ms_per_step": all_ms / a.steps}
            e2e["pipelined"] = {"value": n_all / (piped_ms * 1e-3), "ms_per_step": piped_ms / a.steps, "contexts_per_gpu": a.pipeline}
            if e2e["pipelined"]["value"] > e2e_serial:
                e2e.update({"value": e2e["pipelined"]["value"], "ms_per_step": piped_ms / a.steps, "mode": f"{a.pipeline} contexts per GPU (replicas.Pipeline)"})
        line = {
            "metric": "batch_cnfs_per_s", "value": n_all / (run_ms * 1e-3), "unit": "CNFs/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": run_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "cfg5" + ("" if a.scale == 1.0 else f" x{a.scale}"), "instances": a.batch,
                       "clauses_min": min(weights), "clauses_max": max(weights), "families": "ksat3 / ksat5 / miter / multpar, round robin",
                       "schedule": "static longest-first over ranks (replicas.assign_longest_first)", "flags": "reference defaults",
                       "l2": "each instance is simplified once per step: cold caches", "parallelism": f"instance-parallel x{world}, no collective"},
            "literals_per_s": lit_all / (run_ms * 1e-3),
            "e2e": e2e,
            "gpu_launches": int(sum(t[2] for t in tot)), "clocks": clk,
            "roofline": roofline_block(kstats, peaks, "cfg5"),
        }
        if world == 1 and not a.no_cpu_baseline and os.path.exists(REF_CPU):
            # the reference CPU simplifier on one bounded member of each family; CNFs/s = 1 / mean seconds per instance scaled to the
            # batch's mean instance size (literals), 1 core
            fams = {"cfg1": "ksat3", "cfg2": "ksat5", "cfg3": "miter", "cfg4": "multpar"}
            per = {}
            for wl, nm in fams.items():
                try:
                    path, sV, sC, sL, desc = ref_sample(wl, budget_literals=5_000_000)
                    try:
                        r = run_ref_cpu(path, sL)
                    finally:
                        os.remove(path)
                    per[nm] = {"literals_per_s": sL / (r["ms"] * 1e-3), "sample": desc}
                except Exception as e:  # noqa: BLE001
                    per[nm] = {"error": repr(e)[:120]}
            rates = [v["literals_per_s"] for v in per.values() if "literals_per_s" in v]
            mean_lits = lit_all / max(1.0, n_all)
            line["cpu_baseline"] = {"value": (float(np.mean(rates)) / mean_lits) if rates else None, "unit": "CNFs/s", "cores": 1, "kind": "reference",
                                    "sample": "one 5 M-literal member of each family; CNFs/s = mean literals/s over the families / mean literals per batch instance",
                                    "families": per, "host_cores_available": os.cpu_count()}
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------- our arm
def _pci_path(torch, local):
    p = torch.cuda.get_device_properties(local)
    if hasattr(p, "pci_bus_id"):
        return f"/sys/bus/pci/devices/{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{getattr(p, 'pci_device_id', 0):02x}.0"
    r = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], stdout=subprocess.PIPE, text=True, timeout=20)
    bus = r.stdout.strip().lower()          # 00000000:1b:00.0 -> 0000:1b:00.0
    return "/sys/bus/pci/devices/" + bus[-12:]


def bind_numa(local):
    """Pin this rank to the CPUs (and, by first touch, the memory) of the NUMA node its GPU hangs off, BEFORE any pinned
    buffer is allocated: at N = 8 the e2e step is bound by pinned H2D/D2H through host memory."""
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        path = _pci_path(torch, local)
        node = int(open(path + "/numa_node").read().strip())
        cpulist = open(path + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001 - binding is best effort, the line says what happened
        info["error"] = repr(e)[:120]
    return info


def probe_pcie(torch, local, barrier, mb=256):
    """Pinned host <-> device copy bandwidth of THIS rank while every rank does the same (barrier on both sides): names the
    limiter of the e2e legs - at N = 8 the ranks share the box's host-memory / PCIe root bandwidth."""
    try:
        n = mb << 20
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d_a = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local}")
        d_b = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local}")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        out = {}
        for name, both in (("h2d", (True, False)), ("d2h", (False, True)), ("duplex", (True, True))):
            for rep in range(2):           # first repetition warms up
                barrier()
                t0 = time.perf_counter()
                for _ in range(4):
                    if both[0]:
                        with torch.cuda.stream(s1):
                            d_a.copy_(h_in, non_blocking=True)
                    if both[1]:
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_b, non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            out[name + "_gbs"] = round(4 * n * (both[0] + both[1]) / dt / 1e9, 1)
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:120]}


def offs32_of(offs, alloc):
    """Clause offsets as 32-bit words in pinned memory (sigma_load32: 4 instead of 8 bytes per clause over PCIe) when the
    formula has fewer than 2^32 literals."""
    if int(offs[-1]) >= 1 << 32:
        return offs
    o = alloc(len(offs), np.uint32)
    o[:] = offs
    return o


def kernel_table(kstats, peak, traffic, top=10):
    """[{kernel, ms, launches, share, bytes_per_launch, gbs, frac, traffic}] sorted by time.  bytes = the engine's own
    algorithmic-byte count per kernel name (sigma_kernel_stats; formulas in DESIGN.md 3)."""
    total = sum(v[0] for v in kstats.values()) or 1.0
    rows = []
    for k, (ms, cnt, by) in sorted(kstats.items(), key=lambda kv: -kv[1][0])[:top]:
        gbs = by / (ms * 1e-3) / 1e9 if by > 0 and ms > 0 else None
        rows.append({"kernel": k, "ms": round(ms, 3), "launches": cnt, "share": round(ms / total, 3),
                     "bytes_per_launch": round(by / cnt) if by > 0 else None, "gbs": round(gbs, 1) if gbs else None,
                     "frac": round(gbs / peak, 3) if gbs else None, "traffic": traffic.get(kbase(k))})
    return rows


def roofline_block(kstats, peaks, workload):
    """Roofline of the DOMINANT kernel (largest share of the kernel time of a profiled step)."""
    if not kstats:
        return None
    peak = peaks.get("hbm_gbs")
    src = "measured (MEASURED_PEAKS.json, burst copy bandwidth)"
    if not peak:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = ncu_traffic(workload)
    rows = kernel_table(kstats, peak, traffic)
    d = rows[0]
    return {"bound": "hbm", "kernel": d["kernel"], "launches": d["launches"], "ms_per_launch": d["ms"] / d["launches"],
            "share_of_kernel_time": d["share"], "achieved": d["gbs"], "peak": peak, "peak_source": src, "unit": "GB/s",
            "frac": d["frac"], "algorithmic_bytes_per_launch": d["bytes_per_launch"], "traffic": d["traffic"],
            "bytes_source": "engine-counted algorithmic bytes (sigma_kernel_stats): closed forms for the streaming kernels, "
                            "sum of (4 + 16 + 4|c|) over the clauses of the variables handed to the per-variable kernels",
            "top_kernels": rows}


def measure(a, torch, workload, rank, world, local, barrier, steps, warmup, pipeline, with_clocks=True):
    """One workload on this rank: device-resident steps, e2e steps (serial and through K contexts), a profiled pass."""
    import cnfgen
    from parafrost_b200 import replicas, sigma
    fam, seed, args = workload_spec(workload, rank, a.scale)
    pinned = []

    def alloc(n, dt):
        t = torch.empty(int(n), dtype={np.uint32: torch.int32, np.uint64: torch.int64, np.uint8: torch.uint8}[dt], pin_memory=True)
        pinned.append(t)
        return t.numpy().view(dt)

    V, lits, offs = cnfgen.gen_cnf(fam, seed, args, alloc=alloc)
    offs = offs32_of(offs, alloc)
    C0, L0 = len(offs) - 1, len(lits)
    flags = a.flags.split()
    stream = torch.cuda.Stream()
    s = sigma.Simplifier(local, flags=flags)
    s.set_stream(stream.cuda_stream)
    s.load(V, lits, offs)
    # result buffers for the e2e legs (pinned, sized by the logical capacities of awaken): what newClause() reads
    capC, capL = 2 * C0 + 16, 2 * L0 + 16

    def outbuf():
        return {"bits": alloc(capC, np.uint32), "sizes": alloc(capC, np.uint32), "lits": alloc(capL, np.uint32),
                "eliminated": alloc(V + 1, np.uint8), "resolved": alloc(C0 + L0 + 2, np.uint32), "trail": alloc(3 * (V + 1), np.uint32)}
    out0 = outbuf()

    def step_e2e():
        s.load(V, lits, offs)
        rep = s.simplify()
        return rep, s.store_compact(into=out0)

    for _ in range(warmup):
        s.simplify()
    # ---- value: inputs resident in HBM, no per-kernel events in the timed region
    clocks = ClockSampler(range(world)) if with_clocks and rank == 0 else None   # one sampler for the job's GPUs
    barrier()
    if clocks:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    reps = [s.simplify() for _ in range(steps)]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    rounds = s.rounds()
    # ---- e2e, one context: host buffers in, host buffers out, every step
    for _ in range(min(warmup, 2)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    d2h = 0
    for _ in range(steps):
        rep, st = step_e2e()
        d2h += sum(int(v.nbytes) for v in st.values())
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    # ---- e2e through K contexts (every step still copies its input in and its result out; copies and kernels of
    # neighbouring steps overlap).  All ranks run it together: the PCIe / host-memory contention is part of the number.
    piped = None
    if pipeline > 1:
        bufs = [outbuf() for _ in range(pipeline)]
        with replicas.Pipeline(local, depth=pipeline, flags=flags, compact=True) as pipe:
            pipe.run([(V, lits, offs)] * pipeline, lambda *_: None, bufs)                  # warm-up: arenas, first launches
            torch.cuda.synchronize()
            barrier()
            nsteps = max(steps, 2 * pipeline)
            t0 = time.perf_counter()
            pipe.run([(V, lits, offs)] * nsteps, lambda *_: None, bufs)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            barrier()
            piped = {"ms": (t1 - t0) * 1e3, "steps": nsteps}
    clk = clocks.stop() if clocks else None
    # ---- profiled pass (outside every timed region): CUDA-event pair around each launch + the engine's byte counts
    s.kernel_profile(1)
    psteps = min(steps, 3)
    for _ in range(psteps):
        s.simplify()
    kstats = s.kernel_stats()
    s.kernel_profile(0)
    mem = s.memory()
    s.close()
    return {"fam": fam, "seed": seed, "args": args, "V": V, "C0": C0, "L0": L0, "flags": flags, "ms": ms, "steps": steps, "reps": reps,
            "rounds": rounds, "ms_e2e": ms_e2e, "h2d": int(lits.nbytes + offs.nbytes), "d2h": d2h // steps, "piped": piped, "clocks": clk,
            "kstats": kstats, "psteps": psteps, "memory": mem}


def cpu_baseline_for(workload):
    try:
        path, sV, sC, sL, desc = ref_sample(workload)
        try:
            r = run_ref_cpu(path, sL)
        finally:
            os.remove(path)
        return {"value": sL / (r["ms"] * 1e-3), "unit": UNIT, "cores": 1, "kind": "reference", "sample": desc, "ms": r["ms"],
                "rounds": r["rounds"], "literals_per_round_s": r["literals"] / (r["ms"] * 1e-3), "host_cores_available": os.cpu_count()}
    except Exception as e:  # the baseline is reported, never required for the line
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e!r}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sigma-b200", choices=["sigma-b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=64, help="cfg5: number of instances in the batch")
    ap.add_argument("--pipeline", type=int, default=4,
                    help="K >= 2: e2e steps are also dealt to K engine contexts on the GPU (parafrost_b200.replicas.Pipeline), so that "
                         "copies and kernels of neighbouring steps overlap; 0/1: one context only")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only; the line says so)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg3 block of the default (cfg2) line")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the ref_gpu block (the reference's own GPU build beside the engine)")
    ap.add_argument("--ref-gpu-fixed", action="store_true", help="ref_gpu: also time the reference in its fixed-order mode (single-thread election: slow)")
    ap.add_argument("--flags", default="", help="reference CLI flags for the engine, space separated (e.g. '--phases=5 -no-ere')")
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    if a.impl == "reference-gpu":
        return reference_gpu_arm(a)

    import torch
    import parafrost_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    if not os.path.exists(parafrost_b200.lib_path()):
        parafrost_b200.build()
    torch.cuda.set_device(local)
    numa = bind_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    if a.workload == "cfg5":
        return batch_main(a, torch, dist, barrier, rank, world, local)

    from parafrost_b200 import replicas
    pcie = probe_pcie(torch, local, barrier)
    pcie_sum = None
    if "duplex_gbs" in pcie:
        _, pcie_sum = replicas.reduce_timing(dist, 0.0, float(pcie["duplex_gbs"]), device="cuda")
    m = measure(a, torch, a.workload, rank, world, local, barrier, a.steps, a.warmup, a.pipeline)
    # units: literals of the input formula simplified (one simplify() call = one pass over the formula, however many rounds)
    ms, lit_all = replicas.reduce_timing(dist, m["ms"], float(m["L0"]), device="cuda")      # max over ranks, units summed
    ms_e2e, _ = replicas.reduce_timing(dist, m["ms_e2e"], 0.0, device="cuda")
    ms_pipe = None
    if m["piped"]:
        ms_pipe, _ = replicas.reduce_timing(dist, m["piped"]["ms"] / m["piped"]["steps"], 0.0, device="cuda")   # per step, max over ranks

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        rounds, reps, steps = m["rounds"], m["reps"], m["steps"]
        nrounds = max(1, len(rounds))
        lit_rounds = sum(r["literals_in"] for r in rounds)
        e2e_serial = lit_all * steps / (ms_e2e * 1e-3)
        e2e = {"value": e2e_serial, "unit": UNIT, "ms_per_step": ms_e2e / steps, "mode": "one context: load -> run -> store, nothing overlapped",
               "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
               "api": "sigma_load32 (pinned host CSR, 32-bit offsets) -> sigma_run -> sigma_store_compact (bits, sizes, literals, eliminated, witness stack, trail to pinned host)",
               "pcie_probe": {**pcie, "all_ranks_duplex_gbs": pcie_sum,
                              "note": "pinned host <-> device copies of rank 0 while all ranks copy; a step moves h2d + d2h bytes per rank"}}
        if ms_pipe:
            e2e_pipe = lit_all / (ms_pipe * 1e-3)
            e2e["serial"] = {"value": e2e_serial, "ms_per_step": ms_e2e / steps}
            e2e["pipelined"] = {"value": e2e_pipe, "ms_per_step": ms_pipe, "contexts_per_gpu": a.pipeline, "timer": "host clock around the batch of steps, "
                                "device synchronised on both sides, max over ranks"}
            if e2e_pipe > e2e_serial:
                e2e.update({"value": e2e_pipe, "ms_per_step": ms_pipe,
                            "mode": f"{a.pipeline} contexts per GPU (replicas.Pipeline): every step still copies its input from and its result to pinned "
                                    "host memory; the PCIe legs and kernels of neighbouring steps overlap"})
        line = {
            "metric": METRIC, "value": lit_all * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": a.warmup,
            "ms_per_step": ms / steps, "ms_per_round": ms / steps / nrounds, "rounds_per_step": nrounds,
            "literals_per_round_s": lit_rounds * world * steps / (ms * 1e-3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": a.workload + ("" if a.scale == 1.0 else f" x{a.scale}"), "family": m["fam"], "args": m["args"], "seed": m["seed"],
                       "vars": m["V"], "clauses": m["C0"], "literals": m["L0"], "flags": m["flags"] or "reference defaults (fixed-order election)",
                       "unit_definition": "literals of the input formula per simplify() call (all rounds) / time",
                       "l2": "inputs exceed L2 (no flush needed)" if 4 * m["L0"] > 2 * 126e6 else "inputs fit L2: cold misses only on the first pass of a step",
                       "parallelism": f"replicas x{world} (one CNF per GPU, no collective)", "numa": numa},
            "e2e": e2e,
            "gpu_launches": int(sum(r["kernel_launches"] for r in reps)), "clocks": m["clocks"],
            "result": {"clauses_out": reps[-1]["clauses"], "literals_out": reps[-1]["literals"], "eliminated_vars": reps[-1]["eliminated_vars"],
                       "cnfstate": reps[-1]["cnfstate"], "rounds": [{k: r[k] for k in ("kind", "elected", "eliminated", "resolvents", "clauses", "literals")} for r in rounds]},
            "roofline": roofline_block(m["kstats"], peaks, a.workload),
            "memory": m["memory"],
        }
        if world == 1 and not a.no_cpu_baseline and os.path.exists(REF_CPU):
            line["cpu_baseline"] = cpu_baseline_for(a.workload)
        elif world == 1:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "skipped"}
        # ---- secondary: the workload on which BVE / SUB / GC really fire (cfg2 eliminates nothing: VERDICT r01)
        if world == 1 and a.workload == "cfg2" and not a.no_secondary and a.scale == 1.0:
            try:
                m3 = measure(a, torch, "cfg3", rank, world, local, barrier, min(a.steps, 5), min(a.warmup, 3), a.pipeline, with_clocks=False)
                st3 = m3["steps"]
                sec = {"workload": "cfg3", "vars": m3["V"], "clauses": m3["C0"], "literals": m3["L0"],
                       "value": m3["L0"] * st3 / (m3["ms"] * 1e-3), "unit": UNIT, "ms_per_step": m3["ms"] / st3, "rounds_per_step": len(m3["rounds"]),
                       "ms_per_round": m3["ms"] / st3 / max(1, len(m3["rounds"])),
                       "e2e": {"value": m3["L0"] * st3 / (m3["ms_e2e"] * 1e-3), "ms_per_step": m3["ms_e2e"] / st3, "h2d_bytes_per_step": m3["h2d"],
                               "d2h_bytes_per_step": m3["d2h"],
                               **({"pipelined_ms_per_step": m3["piped"]["ms"] / m3["piped"]["steps"]} if m3["piped"] else {})},
                       "result": {"clauses_out": m3["reps"][-1]["clauses"], "literals_out": m3["reps"][-1]["literals"],
                                  "eliminated_vars": m3["reps"][-1]["eliminated_vars"]},
                       "gpu_launches_per_step": int(m3["reps"][-1]["kernel_launches"]),
                       "roofline": roofline_block(m3["kstats"], peaks, "cfg3")}
                if not a.no_cpu_baseline and os.path.exists(REF_CPU):
                    sec["cpu_baseline"] = cpu_baseline_for("cfg3")
                line["secondary"] = sec
            except Exception as e:  # noqa: BLE001 - the secondary block never costs the headline line
                line["secondary"] = {"workload": "cfg3", "error": repr(e)[:300]}
        if world == 1 and not a.no_ref_gpu and os.path.exists(REF_GPU) and a.scale == 1.0:
            torch.cuda.empty_cache()
            line["ref_gpu"] = ref_gpu_block(a.workload, local, a.ref_gpu_fixed)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


def _only_json_on_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on
    stdout.  Everything but our own print goes to stderr from here on."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _only_json_on_stdout()
    rc = main()
    sys.stdout.flush()
    sys.exit(rc)
