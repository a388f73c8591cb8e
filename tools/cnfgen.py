"""ctypes binding of tools/cnfgen.cpp: the seeded synthetic CNF generators for the BASELINE.json
configs (no network, the reference ships no inputs).  Used by tests/, bench.py and
__graft_entry__.smoke(); not part of the product library."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")

# BASELINE.json configs -> (family, seed, args); sizes recorded in DESIGN.md
CONFIGS = {
    "cfg1": ("ksat", 1, [100_000, 426_000, 3]),             # random 3-SAT n=100k m=426k
    "cfg2": ("ksat", 2, [1_000_000, 21_000_000, 5]),        # random 5-SAT n=1M m=21M
    "cfg3": ("miter", 3, [100_000, 4_950_000, 900, 100, 64]),  # Tseitin AND/XOR miter ~10M vars / ~40M clauses
    "cfg4": ("multpar", 4, [512, 1_000_000]),               # 512x512 array multiplier + parity chain
}


def _newer(target, *srcs):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout)


_gen = None


def build():
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tools", "cnfgen.cpp")
    so = os.path.join(BUILD, "libcnfgen.so")
    if not _newer(so, src):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so + ".tmp%d" % os.getpid(), src])
        os.replace(so + ".tmp%d" % os.getpid(), so)
    exe = os.path.join(BUILD, "cnfgen")
    if not _newer(exe, src):
        _run(["g++", "-O2", "-std=c++17", "-DCNFGEN_MAIN", "-o", exe + ".tmp%d" % os.getpid(), src])
        os.replace(exe + ".tmp%d" % os.getpid(), exe)
    return so


def cnfgen_lib():
    global _gen
    if _gen is None:
        lib = C.CDLL(build())
        lib.cnfgen_create.argtypes = [C.c_char_p, _u64p, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]
        lib.cnfgen_create.restype = C.c_int
        lib.cnfgen_nvars.argtypes = [C.c_void_p]; lib.cnfgen_nvars.restype = C.c_uint32
        lib.cnfgen_nclauses.argtypes = [C.c_void_p]; lib.cnfgen_nclauses.restype = C.c_uint64
        lib.cnfgen_nlits.argtypes = [C.c_void_p]; lib.cnfgen_nlits.restype = C.c_uint64
        lib.cnfgen_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.cnfgen_write_dimacs.argtypes = [C.c_void_p, C.c_char_p]; lib.cnfgen_write_dimacs.restype = C.c_int
        lib.cnfgen_destroy.argtypes = [C.c_void_p]
        _gen = lib
    return _gen


def gen_cnf(family: str, seed: int, args, dimacs_path: str | None = None, alloc=None):
    """-> (max_var, lits uint32[L], offs uint64[C+1]).  `alloc(n, dtype)` may supply the output
    arrays (e.g. views of pinned host memory)."""
    lib = cnfgen_lib()
    h = C.c_void_p()
    a = np.asarray(list(args), np.uint64)
    rc = lib.cnfgen_create(family.encode(), a, len(a), seed, C.byref(h))
    assert rc == 0, f"unknown family {family}"
    try:
        nv, nc, nl = lib.cnfgen_nvars(h), lib.cnfgen_nclauses(h), lib.cnfgen_nlits(h)
        mk = alloc or (lambda n, dt: np.empty(n, dt))
        lits = mk(nl, np.uint32)
        offs = mk(nc + 1, np.uint64)
        lib.cnfgen_copy(h, lits.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p))
        if dimacs_path:
            assert lib.cnfgen_write_dimacs(h, dimacs_path.encode()) == 0
    finally:
        lib.cnfgen_destroy(h)
    return nv, lits, offs
