"""Builds libsigma_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
UNITS = ["scan", "cnf", "otsort", "lcve", "prop", "elim", "api"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def lib_path() -> str:
    return os.path.join(HERE, "libsigma_b200.so")


def _stale(target: str, srcs) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "sigma.h")]
    jobs = []
    for u in UNITS:
        src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(OBJ, u + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout)
        return r.stdout

    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(run, jobs))
    if verbose:
        print("\n".join(outs))
    so = lib_path()
    objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
    if force or jobs or _stale(so, objs):
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs)  # static cudart (nvcc default)
    return so


if __name__ == "__main__":
    print(build(force=True, verbose=True))
