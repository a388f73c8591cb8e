"""N>1 path on CPU: world_size-2 gloo run of the replica plumbing (parafrost_b200/replicas.py) -
the static batch schedule is a disjoint cover, every rank derives the same schedule, timings are
reduced as max-over-ranks and units as a sum.  No collective touches clause data."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from parafrost_b200 import replicas  # noqa: E402


def test_schedule_is_a_disjoint_cover_and_balanced():
    specs = replicas.batch_specs(64)
    w = [replicas.spec_weight(s) for s in specs]
    assert min(w) >= 900_000 and max(w) <= 60_000_000
    for world in (1, 2, 4, 8):
        a = replicas.assign_longest_first(w, world)
        flat = sorted(i for r in a for i in r)
        assert flat == list(range(64))
        loads = [sum(w[i] for i in r) for r in a]
        assert max(loads) - min(loads) <= max(w)      # LPT bound
    assert replicas.assign_longest_first(w, 4) == replicas.assign_longest_first(w, 4)
    assert replicas.rank_seed(2, 3) == 3002


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        specs = replicas.batch_specs(16, scale=1e-3)
        w = [replicas.spec_weight(s) for s in specs]
        mine = replicas.assign_longest_first(w, world)[rank]
        # pretend device times: rank r needs (r+1) ms per instance
        ms, units = replicas.reduce_timing(dist, (rank + 1) * 10.0 * len(mine), float(len(mine)))
        dist.barrier()
        q.put((rank, mine, ms, units))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, ms0, u0), (r1, m1, ms1, u1) = res
    assert sorted(m0 + m1) == list(range(16)) and not set(m0) & set(m1)
    assert ms0 == ms1 == max(10.0 * len(m0), 20.0 * len(m1))     # max over ranks
    assert u0 == u1 == 16.0                                       # units summed
