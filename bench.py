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

    def __init__(self, index):
        self.index, self.p, self.lines = index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
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
def algorithmic_bytes(kernel, C, L, V, E=0):
    """Compulsory HBM bytes of ONE launch (DESIGN.md 'Kernels'): C clause slots, L live literals,
    V variables; 16-byte clause headers, 4-byte literals and occurrence entries."""
    ND = 2 * V + 2
    f = {
        "k_awaken": 8 * C + 4 * L + 16 * C + 4 * L + 16 * C + 4 * ND,   # prep + the fused first-round histogram and sort keys
        "k_hist_key": 16 * C + 4 * L + 16 * C + 4 * ND,
        "k_hist": 16 * C + 4 * L + 4 * ND,
        "k_ot_part": 16 * C + 4 * L + 8 * L,
        "k_ot_place": 8 * L + 4 * L + 12 * ND,
        "k_sort_small": 4 * L + 16 * L + 4 * L + 8 * ND,
        "k_sort_med": 4 * L + 16 * L + 4 * L,
        "k_sort_lists": 4 * L + 16 * L + 4 * L + 8 * ND,
        "k_count": 16 * C,
        "k_gc_copy": 16 * C + 4 * L + 8 * C + 16 * C + 4 * L,
        "k_gc_flags": 16 * C + 8 * C,
        "k_store_arrays": 16 * C + 4 * L + 8 * C + 16 * C + 4 * L,
    }
    return f.get(kernel)


def ncu_traffic(workload):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of this workload's kernels
    from the newest committed `ncu --set full` capture (profiles/*_ncu_traffic_<workload>.json,
    written by tools/ncu_digest.py --json)."""
    import glob
    best = {}
    def version(f):   # ..._vNN.json, natural order
        m = re.search(r"_v(\d+)[a-z]*\.json$", f)
        return int(m.group(1)) if m else -1
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_ncu_traffic_{workload}*.json")), key=version):
        try:
            best.update(json.load(open(f)).get("kernels", {}))
        except (OSError, ValueError):
            pass
    return best


def kbase(name):
    """'(k_ot_part<3, 5>)' / 'void k_sub<4>' -> 'k_ot_part' / 'k_sub' (the LAUNCH macro stringifies its argument)."""
    return name.replace("void ", "").strip("() ").split("<")[0]


def roofline(ktimes, C, L, V, peaks, workload="cfg2"):
    """Roofline of the dominant kernel.  Streaming kernels have closed-form algorithmic bytes
    (DESIGN.md 3); the per-variable kernels (MIS rounds, BVE, SUB, ERE) are latency / instruction
    bound gather kernels whose bytes depend on the elected set - when one of those leads, the line
    names it (`dominant`) and carries the roofline of the largest kernel that HAS a byte formula."""
    if not ktimes:
        return None
    peak = peaks.get("hbm_gbs")
    src = "measured (MEASURED_PEAKS.json)"
    if not peak:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    total = sum(v[0] for v in ktimes.values())
    order = sorted(ktimes.items(), key=lambda kv: -kv[1][0])
    dom_name, (dom_ms, dom_cnt) = order[0]
    name, (ms, cnt) = next(((k, v) for k, v in order if algorithmic_bytes(kbase(k), C, L, V)), order[0])
    traffic = ncu_traffic(workload)
    b = algorithmic_bytes(kbase(name), C, L, V)
    out = {"bound": "hbm", "kernel": name, "launches": cnt, "ms_per_launch": ms / cnt, "share_of_kernel_time": ms / total if total else None,
           "peak": peak, "peak_source": src, "unit": "GB/s", "traffic": traffic.get(kbase(name))}
    if b:
        ach = b / (ms / cnt * 1e-3) / 1e9
        out.update({"achieved": ach, "frac": ach / peak, "algorithmic_bytes_per_launch": b})
    else:
        out.update({"achieved": None, "frac": None, "algorithmic_bytes_per_launch": None})
    if dom_name != name:
        out["dominant"] = {"kernel": dom_name, "ms_per_launch": dom_ms / dom_cnt, "launches": dom_cnt, "share_of_kernel_time": dom_ms / total,
                           "traffic": traffic.get(kbase(dom_name)),
                           "note": "per-variable gather kernel: latency/instruction bound, no closed-form bytes (DESIGN.md 3)"}
    out["top_kernels"] = [{"kernel": k, "ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / total, 3),
                           "gbs": (lambda bb: round(bb / (v[0] / v[1] * 1e-3) / 1e9, 1) if bb else None)(algorithmic_bytes(kbase(k), C, L, V))}
                          for k, v in order[:8]]
    return out


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
    lit = sum(r["literals"] for r in runs)
    value = lit / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "ms_per_round": ms / max(1, sum(r["rounds"] for r in runs)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": a.workload, "sample": desc, "solver": "reference CPU v3.2.5 (src/cpu), -no-solve -profilesimp, default inprocessing"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": desc,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
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
        inst.append((V, lits, offs, keep))
    maxC = max((len(x[2]) - 1 for x in inst), default=1)
    maxL = max((len(x[1]) for x in inst), default=1)
    maxV = max((x[0] for x in inst), default=1)
    pin = []

    def palloc(n, dt):
        t = torch.empty(int(n), dtype={np.uint32: torch.int32, np.uint64: torch.int64}[dt], pin_memory=True)
        pin.append(t)
        return t.numpy().view(dt)
    outbuf = {"bits": palloc(2 * maxC + 16, np.uint32), "sig": palloc(2 * maxC + 16, np.uint32), "offs": palloc(2 * maxC + 17, np.uint64),
              "lits": palloc(2 * maxL + 16, np.uint32), "eliminated": np.zeros(maxV + 1, np.uint8), "resolved": palloc(maxC + maxL + 2, np.uint32),
              "trail": palloc(3 * (maxV + 1), np.uint32)}

    def one_pass(timed):
        """-> (ms of sigma_run summed over my instances, ms of the whole pass, launches, literals, h2d, d2h)"""
        run_ms, launches, lit, h2d, d2h = 0.0, 0, 0, 0, 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for V, lits, offs, _ in inst:
            s.load(V, lits, offs)
            rep = s.simplify()
            outbuf["eliminated"] = np.zeros(V + 1, np.uint8)
            st = s.store(into=outbuf)
            run_ms += rep["ms_device"]; launches += rep["kernel_launches"]
            lit += sum(r["literals_in"] for r in s.rounds())
            h2d += int(lits.nbytes + offs.nbytes); d2h += sum(int(v.nbytes) for v in st.values())
        e1.record(stream)
        stream.synchronize()
        return run_ms, e0.elapsed_time(e1), launches, lit, h2d, d2h

    for _ in range(a.warmup):
        one_pass(False)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    tot = [one_pass(True) for _ in range(a.steps)]
    barrier()
    clk = clocks.stop()
    # ---- optional: the same passes through K contexts per GPU (replicas.Pipeline): copies of one instance overlap the
    # kernels of another; every instance is still loaded from and stored to pinned host memory
    piped_ms = None
    if a.pipeline > 1:
        bufs = [{"bits": palloc(2 * maxC + 16, np.uint32), "sig": palloc(2 * maxC + 16, np.uint32), "offs": palloc(2 * maxC + 17, np.uint64),
                 "lits": palloc(2 * maxL + 16, np.uint32), "eliminated": np.zeros(maxV + 1, np.uint8),
                 "resolved": palloc(maxC + maxL + 2, np.uint32), "trail": palloc(3 * (maxV + 1), np.uint32)} for _ in range(a.pipeline)]
        jobs = [(V, lits, offs) for V, lits, offs, _ in inst]
        with replicas.Pipeline(local, depth=a.pipeline) as pipe:
            pipe.run(jobs, lambda *_: None, bufs)          # warm-up
            torch.cuda.synchronize()
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(a.steps):
                pipe.run(jobs, lambda *_: None, bufs)
            torch.cuda.synchronize()
            p1.record()
            p1.synchronize()
            piped_ms = p0.elapsed_time(p1)
        piped_ms, _ = replicas.reduce_timing(dist, piped_ms, 0.0, device="cuda")
    run_ms = sum(t[0] for t in tot); all_ms = sum(t[1] for t in tot)
    n_mine = len(inst) * a.steps
    run_ms, n_all = replicas.reduce_timing(dist, run_ms, float(n_mine), device="cuda")
    all_ms, lit_all = replicas.reduce_timing(dist, all_ms, float(sum(t[3] for t in tot)), device="cuda")
    if rank == 0:
        line = {
            "metric": "batch_cnfs_per_s", "value": n_all / (run_ms * 1e-3), "unit": "CNFs/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": run_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "cfg5" + ("" if a.scale == 1.0 else f" x{a.scale}"), "instances": a.batch,
                       "clauses_min": min(weights), "clauses_max": max(weights), "families": "ksat3 / ksat5 / miter / multpar, round robin",
                       "schedule": "static longest-first over ranks (replicas.assign_longest_first)", "flags": "reference defaults",
                       "l2": "each instance is simplified once per step: cold caches", "parallelism": f"instance-parallel x{world}, no collective"},
            "literals_per_s": lit_all / (run_ms * 1e-3),
            "e2e": {"value": n_all / (all_ms * 1e-3), "unit": "CNFs/s", "ms_per_step": all_ms / a.steps,
                    "h2d_bytes_per_step": tot[-1][4], "d2h_bytes_per_step": tot[-1][5]},
            **({"e2e_pipelined": {"value": n_all / (piped_ms * 1e-3), "unit": "CNFs/s", "contexts_per_gpu": a.pipeline,
                                  "ms_per_step": piped_ms / a.steps}} if piped_ms else {}),
            "gpu_launches": int(sum(t[2] for t in tot)), "clocks": clk,
            "roofline": None, "cpu_baseline": {"value": None, "unit": "CNFs/s", "cores": 0, "kind": "reference", "sample": "not run for the batch workload (see cfg2)"},
        }
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sigma-b200", choices=["sigma-b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=64, help="cfg5: number of instances in the batch")
    ap.add_argument("--pipeline", type=int, default=0,
                    help="K >= 2: additionally report e2e_pipelined - the e2e steps dealt to K engine contexts on the GPU "
                         "(parafrost_b200.replicas.Pipeline), so that copies and kernels of neighbouring steps overlap")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only; the line says so)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", default="", help="reference CLI flags for the engine, space separated (e.g. '--phases=5 -no-ere')")
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import cnfgen
    import parafrost_b200
    from parafrost_b200 import sigma

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    if not os.path.exists(parafrost_b200.lib_path()):
        parafrost_b200.build()
    torch.cuda.set_device(local)
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

    # ---- synthetic input of the named shape, in pinned host memory
    fam, seed, args = workload_spec(a.workload, rank, a.scale)
    pinned = []

    def alloc(n, dt):
        t = torch.empty(int(n), dtype={np.uint32: torch.int32, np.uint64: torch.int64}[dt], pin_memory=True)
        pinned.append(t)
        return t.numpy().view(dt)

    V, lits, offs = cnfgen.gen_cnf(fam, seed, args, alloc=alloc)
    C0, L0 = len(offs) - 1, len(lits)
    flags = a.flags.split()
    stream = torch.cuda.Stream()
    s = sigma.Simplifier(local, flags=flags)
    s.set_stream(stream.cuda_stream)
    s.load(V, lits, offs)
    # result buffers for the e2e leg (pinned, sized by the logical capacities of awaken)
    capC, capL = 2 * C0 + 16, 2 * L0 + 16
    outbuf = {"bits": alloc(capC, np.uint32), "sig": alloc(capC, np.uint32), "offs": alloc(capC + 1, np.uint64),
              "lits": alloc(capL, np.uint32), "eliminated": np.zeros(V + 1, np.uint8), "resolved": alloc(C0 + L0 + 2, np.uint32),
              "trail": alloc(3 * (V + 1), np.uint32)}

    def step_resident():
        return s.simplify()

    def step_e2e():
        s.load(V, lits, offs)
        rep = s.simplify()
        st = s.store(into=outbuf)
        return rep, st

    for _ in range(a.warmup):
        step_resident()
    # ---- value: inputs resident in HBM
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    s.kernel_profile(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    reps = [step_resident() for _ in range(a.steps)]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    ktimes = s.kernel_times()
    s.kernel_profile(0)
    rounds = s.rounds()
    # ---- e2e: host buffers in, host buffers out
    for _ in range(min(a.warmup, 2)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    d2h = 0
    for _ in range(a.steps):
        rep, st = step_e2e()
        d2h += sum(int(v.nbytes) for v in st.values())
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clk = clocks.stop()
    # ---- optional: the same e2e steps through K contexts (every step still copies its input in and its result out)
    piped = None
    if a.pipeline > 1:
        from parafrost_b200 import replicas as _rep
        bufs = [{"bits": alloc(capC, np.uint32), "sig": alloc(capC, np.uint32), "offs": alloc(capC + 1, np.uint64),
                 "lits": alloc(capL, np.uint32), "eliminated": np.zeros(V + 1, np.uint8), "resolved": alloc(C0 + L0 + 2, np.uint32),
                 "trail": alloc(3 * (V + 1), np.uint32)} for _ in range(a.pipeline)]
        with _rep.Pipeline(local, depth=a.pipeline, flags=flags) as pipe:
            pipe.run([(V, lits, offs)] * a.pipeline, lambda *_: None, bufs)                  # warm-up: arenas, first launches
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            nsteps = max(a.steps, a.pipeline)
            pipe.run([(V, lits, offs)] * nsteps, lambda *_: None, bufs)
            torch.cuda.synchronize()
            p1.record()
            p1.synchronize()
            piped = {"ms": p0.elapsed_time(p1), "steps": nsteps}

    lit_step = sum(r["literals_in"] for r in rounds)
    nrounds = max(1, len(rounds))
    launches = sum(r["kernel_launches"] for r in reps)
    from parafrost_b200 import replicas
    ms, lit_all = replicas.reduce_timing(dist, ms, float(lit_step), device="cuda")      # max over ranks, units summed
    ms_e2e, _ = replicas.reduce_timing(dist, ms_e2e, 0.0, device="cuda")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        meanC = float(np.mean([r["clauses"] for r in rounds])) if rounds else C0
        meanL = float(np.mean([r["literals_in"] for r in rounds])) if rounds else L0
        line = {
            "metric": METRIC, "value": lit_all * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "ms_per_round": ms / a.steps / nrounds, "rounds_per_step": nrounds,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": a.workload + ("" if a.scale == 1.0 else f" x{a.scale}"), "family": fam, "args": args, "seed": seed,
                       "vars": V, "clauses": C0, "literals": L0, "flags": flags or "reference defaults (fixed-order election)",
                       "l2": "inputs exceed L2 (no flush needed)" if 4 * L0 > 2 * 126e6 else "inputs fit L2: cold misses only on the first pass of a step",
                       "parallelism": f"replicas x{world} (one CNF per GPU, no collective)"},
            "e2e": {"value": lit_all * a.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(lits.nbytes + offs.nbytes), "d2h_bytes_per_step": d2h // a.steps},
            "gpu_launches": int(launches), "clocks": clk,
            **({"e2e_pipelined": {"value": lit_step * piped["steps"] / (piped["ms"] * 1e-3), "unit": UNIT, "contexts": a.pipeline,
                                  "steps": piped["steps"], "ms_per_step": piped["ms"] / piped["steps"],
                                  "note": "rank 0 only; every step copies its input from and its result to pinned host memory, "
                                          "steps run on K contexts so that PCIe legs and kernels of neighbouring steps overlap"}}
               if piped else {}),
            "result": {"clauses_out": reps[-1]["clauses"], "literals_out": reps[-1]["literals"], "eliminated_vars": reps[-1]["eliminated_vars"],
                       "cnfstate": reps[-1]["cnfstate"]},
            "roofline": roofline(ktimes, meanC, meanL, V, peaks, a.workload),
        }
        if world == 1 and not a.no_cpu_baseline and os.path.exists(REF_CPU):
            try:
                path, sV, sC, sL, desc = ref_sample(a.workload)
                try:
                    r = run_ref_cpu(path, sL)
                finally:
                    os.remove(path)
                line["cpu_baseline"] = {"value": r["literals"] / (r["ms"] * 1e-3), "unit": UNIT, "cores": 1, "kind": "reference",
                                        "sample": desc, "ms": r["ms"], "rounds": r["rounds"], "host_cores_available": os.cpu_count()}
            except Exception as e:  # the baseline is reported, never required for the line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e!r}"[:200]}
        elif world == 1:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "skipped"}
        print(json.dumps(line))
    s.close()
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
