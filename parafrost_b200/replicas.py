"""Instance-parallel plumbing ("replicas only").

A CNF does not shard - BVE/SUB/ERE need one occurrence table, one independent set and one clause
arena (the reference never touches a second GPU, src/gpu/constants.cuh:27 MASTER_GPU) - so N GPUs
simplify N independent formulas, one engine context per formula, and nothing crosses NVLink on the
data path.  This module holds the host logic around that: which rank takes which instance of a
batch (BASELINE.json config 5: 64 mixed CNFs over 1/2/4/8 GPUs) and how per-rank device timings are
combined (max over ranks, units summed).  torch.distributed is used for the barrier and the two
reductions only; it works the same on gloo (CPU tests) and nccl (GPU box)."""
from __future__ import annotations

# BASELINE.json config 5: 64 mixed synthetic CNFs of 1-50 M clauses.  (family, seed, args); the
# clause counts follow a fixed geometric ladder so the batch is identical on every run.
def batch_specs(n: int = 64, scale: float = 1.0):
    specs = []
    for i in range(n):
        fam = ("ksat3", "ksat5", "miter", "multpar")[i % 4]
        # 1 M .. 50 M clauses, geometric in i
        target = int(1_000_000 * (50.0 ** (i / max(1, n - 1))) * scale)
        target = max(target, 64)
        if fam == "ksat3":
            specs.append(("ksat", 500 + i, [max(8, int(target / 4.26)), target, 3]))
        elif fam == "ksat5":
            specs.append(("ksat", 500 + i, [max(8, target // 21), target, 5]))
        elif fam == "miter":
            # cfg3's generator: ~8 clauses per gate pair; inputs scale with the square root
            gates = max(8, target // 8)
            specs.append(("miter", 500 + i, [max(4, int(gates ** 0.5)), gates, 900, 100, 64]))
        else:
            # n x n array multiplier has ~17 n^2 clauses, the parity chain 4 per link: half / half
            nbits = max(4, int((target / 2 / 17.0) ** 0.5))
            specs.append(("multpar", 500 + i, [nbits, max(8, target // 8)]))
    return specs


def spec_weight(spec) -> int:
    """Clause-count estimate used for the static schedule (no generation needed)."""
    fam, _, a = spec
    if fam == "ksat":
        return int(a[1])
    if fam == "miter":
        return int(a[1]) * 8
    if fam == "multpar":
        return int(17 * a[0] * a[0] + 4 * a[1])
    return 1


def assign_longest_first(weights, world: int):
    """Static longest-processing-time-first schedule: returns world lists of instance indices.
    Deterministic (ties by index), every instance exactly once."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += weights[i]
    return out


def reduce_timing(dist, ms_local: float, units_local: float, device="cpu"):
    """(max over ranks of ms, sum over ranks of units).  `dist` is torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms_local), float(units_local)
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def rank_seed(seed: int, rank: int) -> int:
    """Replica r of a single-config bench simplifies its own formula of the same shape."""
    return seed + 1000 * rank


class Pipeline:
    """Several engine contexts on ONE device for a stream of independent instances (BASELINE.json config 5 on each
    rank): every context has its own CUDA stream (sigma_create) and its own host thread, and each thread takes the
    next instance off a shared queue and runs load -> simplify -> store on it.  With `depth` >= 3 the host->device
    copy of instance i+1, the kernels of instance i and the device->host copy of instance i-1 are in flight at the
    same time (two copy engines + the SMs), where a single context leaves the GPU idle during the PCIe legs - which
    are most of an end-to-end step (24.5 of 30 ms on cfg2).  The reference has nothing like it (one
    Solver per process, blocking copies); nothing crosses between the contexts, results are bit-identical to the
    sequential run.  The C ABI calls release the GIL (ctypes), so plain threads are enough.

    make() -> an object with load / simplify / rounds / store / close (sigma.Simplifier); injectable for CPU tests.
    """

    def __init__(self, device: int = 0, depth: int = 3, flags=(), make=None, compact: bool = False, **opts):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self._compact = compact   # results leave through sigma_store_compact (bits, sizes, literals): what newClause() reads
        if make is None:
            from . import sigma

            def make():
                return sigma.Simplifier(device, flags=flags, **opts)
        self._ctx = []
        try:
            for _ in range(depth):
                self._ctx.append(make())
        except Exception:
            self.close()
            raise

    @property
    def depth(self) -> int:
        return len(self._ctx)

    def run(self, jobs, consume, outbufs=None):
        """jobs: sequence of (max_var, lits, offs) or (max_var, lits, offs, meta) in HOST memory (pinned for real
        overlap).  consume(i, report, rounds, stored) is called on the worker thread that finished instance i;
        `stored` are views into that worker's output buffers (outbufs[w], a dict for Simplifier.store(into=...), or
        fresh arrays when outbufs is None) and are only valid during the call.  Instances are started in index
        order; the first exception stops the queue and is re-raised here."""
        import threading
        jobs = list(jobs)
        if outbufs is not None and len(outbufs) < len(self._ctx):
            raise ValueError("one output buffer set per context is needed")
        lock = threading.Lock()
        state = {"next": 0, "error": None}

        def worker(w):
            s = self._ctx[w]
            while True:
                with lock:
                    if state["error"] is not None or state["next"] >= len(jobs):
                        return
                    i = state["next"]
                    state["next"] += 1
                try:
                    job = jobs[i]
                    s.load(job[0], job[1], job[2], meta=job[3] if len(job) > 3 else None)
                    rep = s.simplify()
                    store = s.store_compact if self._compact else s.store
                    stored = store(into=outbufs[w]) if outbufs is not None else store()
                    consume(i, rep, s.rounds(), stored)
                except BaseException as e:   # noqa: BLE001 - handed to the caller
                    with lock:
                        if state["error"] is None:
                            state["error"] = e
                    return

        threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(len(self._ctx))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if state["error"] is not None:
            raise state["error"]

    def close(self):
        for s in self._ctx:
            try:
                s.close()
            except Exception:
                pass
        self._ctx = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
