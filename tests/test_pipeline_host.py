"""Host logic of parafrost_b200.replicas.Pipeline (several contexts on one device, one thread each) with a fake
engine: every instance exactly once, started in order, contexts really run concurrently, errors surface."""
import threading
import time

import numpy as np
import pytest

from parafrost_b200 import replicas


class FakeSimplifier:
    live = 0
    peak = 0
    lock = threading.Lock()

    def __init__(self, log, fail_on=None):
        self.log, self.fail_on, self.closed, self.cur = log, fail_on, False, None

    def load(self, V, lits, offs, meta=None):
        with FakeSimplifier.lock:
            FakeSimplifier.live += 1
            FakeSimplifier.peak = max(FakeSimplifier.peak, FakeSimplifier.live)
        self.cur = (V, int(lits.sum()), len(offs) - 1, meta is not None)
        self.log.append(("load", V))
        time.sleep(0.01)

    def simplify(self):
        if self.fail_on is not None and self.cur[0] == self.fail_on:
            with FakeSimplifier.lock:
                FakeSimplifier.live -= 1
            raise RuntimeError(f"boom on {self.cur[0]}")
        time.sleep(0.02)
        return {"ms_device": 1.0, "kernel_launches": 7, "V": self.cur[0]}

    def rounds(self):
        return [{"literals_in": self.cur[1]}]

    def store(self, into=None):
        time.sleep(0.01)
        with FakeSimplifier.lock:
            FakeSimplifier.live -= 1
        out = {"sum": np.array([self.cur[1]], np.uint64)}
        if into is not None:
            into["sum"][0] = self.cur[1]
            return {"sum": into["sum"][:1]}
        return out

    def close(self):
        self.closed = True


def make_jobs(n):
    rng = np.random.default_rng(3)
    jobs = []
    for i in range(n):
        lits = rng.integers(2, 100, size=10 + i).astype(np.uint32)
        offs = np.array([0, len(lits)], np.uint64)
        jobs.append((i + 1, lits, offs) if i % 2 else (i + 1, lits, offs, np.zeros(1, np.uint32)))
    return jobs


@pytest.mark.parametrize("depth", [1, 3])
def test_every_instance_once_and_in_flight_together(depth):
    FakeSimplifier.live = FakeSimplifier.peak = 0
    log, made = [], []

    def make():
        made.append(FakeSimplifier(log))
        return made[-1]
    jobs = make_jobs(12)
    got = {}
    outbufs = [{"sum": np.zeros(1, np.uint64)} for _ in range(depth)]
    with replicas.Pipeline(depth=depth, make=make) as p:
        assert p.depth == depth
        p.run(jobs, lambda i, rep, rounds, st: got.__setitem__(i, (rep["V"], int(st["sum"][0]), rounds[0]["literals_in"])), outbufs)
    assert sorted(got) == list(range(12))
    for i, (V, ssum, lit) in got.items():
        assert V == i + 1 and ssum == int(jobs[i][1].sum()) == lit
    loads = [v for k, v in log if k == "load"]
    assert sorted(loads) == list(range(1, 13))
    assert all(abs(loads[k] - (k + 1)) < depth + 1 for k in range(12))   # started in index order, up to the races of `depth` threads
    assert FakeSimplifier.peak == min(depth, 12) if depth == 1 else FakeSimplifier.peak >= 2
    assert all(s.closed for s in made)


def test_first_error_stops_the_queue_and_is_raised():
    FakeSimplifier.live = FakeSimplifier.peak = 0
    log = []
    p = replicas.Pipeline(depth=2, make=lambda: FakeSimplifier(log, fail_on=4))
    done = []
    with pytest.raises(RuntimeError, match="boom on 4"):
        p.run(make_jobs(40), lambda i, *_: done.append(i))
    p.close()
    assert 3 not in done and len(done) < 40


def test_needs_one_output_buffer_per_context():
    p = replicas.Pipeline(depth=3, make=lambda: FakeSimplifier([]))
    with pytest.raises(ValueError):
        p.run(make_jobs(2), lambda *a: None, outbufs=[{}])
    p.close()
    with pytest.raises(ValueError):
        replicas.Pipeline(depth=0, make=lambda: FakeSimplifier([]))
