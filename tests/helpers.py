"""ctypes bindings for the test-side native helpers (generator + CPU oracle).

Nothing in here is imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import sgd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _newer(target, *srcs):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout)


# ------------------------------------------------------------------ generator (tools/cnfgen.py)
import sys
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cnfgen import CONFIGS, cnfgen_lib, gen_cnf  # noqa: E402,F401


# ------------------------------------------------------------------ oracle
class OracleOpts(C.Structure):
    _fields_ = [
        ("phases", C.c_int32), ("ve_en", C.c_int32), ("ve_plus_en", C.c_int32), ("sub_en", C.c_int32),
        ("bce_en", C.c_int32), ("ere_en", C.c_int32), ("all_en", C.c_int32),
        ("mu_pos", C.c_uint32), ("mu_neg", C.c_uint32), ("lcve_min_vars", C.c_uint32),
        ("lcve_max_occurs", C.c_uint32), ("lcve_clause_max", C.c_int32), ("phase_lits_min", C.c_int32),
        ("shrink_rate", C.c_int32), ("lits_mul", C.c_double),
        ("ve_fun_en", C.c_int32), ("ve_lbound_en", C.c_int32), ("ve_clause_max", C.c_uint32),
        ("xor_max_arity", C.c_uint32), ("ere_clause_max", C.c_int32), ("ere_max_occurs", C.c_uint32),
        ("sub_max_occurs", C.c_uint32), ("bce_max_occurs", C.c_uint32), ("sh_max_bve_out1", C.c_uint32),
        ("sigma_calls", C.c_int32), ("final_gc", C.c_int32), ("aggr_cnf_sort", C.c_int32), ("lcve_fast", C.c_int32),
    ]


_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        src = os.path.join(ROOT, "oracle", "sigma_oracle.cpp")
        hdr = os.path.join(ROOT, "oracle", "sigma_oracle.h")
        so = os.path.join(ROOT, "oracle", "libsigma_oracle.so")
        if not _newer(so, src, hdr):
            _run(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile"), so])
        lib = C.CDLL(so)
        P = C.c_void_p
        lib.oracle_default_opts.argtypes = [C.POINTER(OracleOpts)]
        lib.oracle_normalize_opts.argtypes = [C.POINTER(OracleOpts)]
        lib.oracle_create.argtypes = [C.POINTER(OracleOpts), C.c_uint32, C.c_uint64, _u32p, _u64p, P, P, P, C.POINTER(P)]
        lib.oracle_create.restype = C.c_int
        lib.oracle_run.argtypes = [P]; lib.oracle_run.restype = C.c_int
        lib.oracle_rounds.argtypes = [P]; lib.oracle_rounds.restype = C.c_int
        lib.oracle_round_stats.argtypes = [P, _u64p]
        lib.oracle_record_ends.argtypes = [P, C.c_uint64, P]; lib.oracle_record_ends.restype = C.c_uint64
        lib.oracle_hash_segments.argtypes = [P, P, P, C.c_uint64, P]; lib.oracle_hash_segments.restype = None
        for f in ("oracle_num_clauses", "oracle_num_literals", "oracle_num_resolved", "oracle_num_trail"):
            getattr(lib, f).argtypes = [P]; getattr(lib, f).restype = C.c_uint64
        lib.oracle_copy_result.argtypes = [P, _u32p, _u32p, _u64p, _u32p, _u8p, _u32p, _u32p]
        lib.oracle_keep_snapshots.argtypes = [P, C.c_int]
        lib.oracle_set_assumed.argtypes = [P, P]
        lib.oracle_snapshot_clauses.argtypes = [P, C.c_int]; lib.oracle_snapshot_clauses.restype = C.c_uint64
        lib.oracle_snapshot_literals.argtypes = [P, C.c_int]; lib.oracle_snapshot_literals.restype = C.c_uint64
        lib.oracle_copy_snapshot.argtypes = [P, C.c_int, _u32p, _u32p, _u64p, _u32p]
        lib.oracle_destroy.argtypes = [P]
        lib.oracle_num_elections.argtypes = [P]; lib.oracle_num_elections.restype = C.c_int
        lib.oracle_election_size.argtypes = [P, C.c_int]; lib.oracle_election_size.restype = C.c_uint64
        lib.oracle_copy_election.argtypes = [P, C.c_int, _u32p]
        lib.oracle_enable_proof.argtypes = [P, C.c_int]
        lib.oracle_proof_chunks.argtypes = [P]; lib.oracle_proof_chunks.restype = C.c_int
        lib.oracle_proof_chunk_size.argtypes = [P, C.c_int]; lib.oracle_proof_chunk_size.restype = C.c_uint64
        lib.oracle_copy_proof_chunk.argtypes = [P, C.c_int, _u8p]
        lib.oracle_proof_capacity.argtypes = [P]; lib.oracle_proof_capacity.restype = C.c_uint32
        lib.oracle_prep.argtypes = [C.c_uint64, _u32p, _u64p, _u32p]
        lib.oracle_histogram.argtypes = [C.c_uint64, _u32p, C.c_uint32, _u32p]
        lib.oracle_extend_model.argtypes = [_u8p, C.c_uint32, _u32p, C.c_uint64]; lib.oracle_extend_model.restype = C.c_uint64
        lib.oracle_check_model.argtypes = [_u8p, C.c_uint64, _u32p, _u64p]; lib.oracle_check_model.restype = C.c_uint64
        _orc = lib
    return _orc


FLAG_MAP = {  # reference CLI flag -> option override (src/gpu/options.cpp:24-43, options.cu:36-60)
    "-no-ere": {"ere_en": 0}, "-ere": {"ere_en": 1}, "-no-vefunction": {"ve_fun_en": 0}, "-bce": {"bce_en": 1}, "-all": {"all_en": 1},
    "-no-sub": {"sub_en": 0}, "-no-veextend": {"ve_plus_en": 0}, "-no-ve": {"ve_en": 0}, "-velitsbound": {"ve_lbound_en": 1}, "-aggresivesort": {"aggr_cnf_sort": 1},
    "-no-lcvefast": {"lcve_fast": 0}, "-lcvefast": {"lcve_fast": 1}, "-quiet": {},
}
VALUE_FLAGS = {
    "--phases": "phases", "--mupos": "mu_pos", "--muneg": "mu_neg", "--electionsmin": "lcve_min_vars",
    "--electionsmax": "lcve_max_occurs", "--lcveclausemax": "lcve_clause_max", "--eliminatedlitsmin": "phase_lits_min",
    "--collectfreq": "shrink_rate", "--literalsmul": "lits_mul", "--resolventmax": "ve_clause_max",
    "--xormaxarity": "xor_max_arity", "--ereclausemax": "ere_clause_max", "--eremaxoccurs": "ere_max_occurs",
    "--submaxoccurs": "sub_max_occurs", "--bcemaxoccurs": "bce_max_occurs",
}


def opts_from_flags(flags) -> dict:
    o = {}
    for f in flags:
        if "=" in f:
            k, v = f.split("=", 1)
            if k in VALUE_FLAGS:
                o[VALUE_FLAGS[k]] = float(v) if k == "--literalsmul" else int(v)
            elif k not in ("--mapperc", "--ereminthreads"):   # host-side / launch-shape options: no effect on the result
                raise ValueError(f"unknown flag {f}")
        else:
            o.update(FLAG_MAP[f])
    return o


def make_oracle_opts(**over) -> OracleOpts:
    lib = oracle_lib()
    o = OracleOpts()
    lib.oracle_default_opts(C.byref(o))
    for k, v in over.items():
        setattr(o, k, v)
    lib.oracle_normalize_opts(C.byref(o))
    return o


def run_oracle(max_var, lits, offs, meta=None, snapshots=False, vorg=None, vstate=None, assumed=None, proof=False, **over):
    """Run the CPU oracle -> (Dump, round_stats uint64[R,5], [snapshot Dumps]).
    proof=True: the device DRAT stream is on; Dump.extra["proof"] = list of chunks (bytes), extra["proof_cap"]."""
    lib = oracle_lib()
    o = make_oracle_opts(**over)
    h = C.c_void_p()
    lits = np.ascontiguousarray(lits, np.uint32)
    offs = np.ascontiguousarray(offs, np.uint64)
    meta_p = None
    if meta is not None:
        meta = np.ascontiguousarray(meta, np.uint32)
        meta_p = meta.ctypes.data_as(C.c_void_p)
    keep = [None if a is None else np.ascontiguousarray(a, dt) for a, dt in ((vorg, np.uint32), (vstate, np.uint8), (assumed, np.uint8))]
    ptr = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in keep]
    lib.oracle_create(C.byref(o), max_var, len(offs) - 1, lits, offs, meta_p, ptr[0], ptr[1], C.byref(h))
    try:
        lib.oracle_keep_snapshots(h, int(snapshots))
        lib.oracle_enable_proof(h, int(proof))
        if keep[2] is not None:
            lib.oracle_set_assumed(h, ptr[2])
        state = lib.oracle_run(h)
        nc, nl = lib.oracle_num_clauses(h), lib.oracle_num_literals(h)
        nr, nt = lib.oracle_num_resolved(h), lib.oracle_num_trail(h)
        bits = np.empty(nc, np.uint32); sig = np.empty(nc, np.uint32)
        o_offs = np.empty(nc + 1, np.uint64); o_lits = np.empty(nl, np.uint32)
        elim = np.empty(max_var + 1, np.uint8)
        res = np.empty(nr, np.uint32); trail = np.empty(nt, np.uint32)
        lib.oracle_copy_result(h, bits, sig, o_offs, o_lits, elim, res, trail)
        d = sgd.Dump.from_arrays(max_var, state, bits, sig, o_offs, o_lits, elim, res, trail)
        R = lib.oracle_rounds(h)
        rs = np.zeros((R, 5), np.uint64)
        if R:
            lib.oracle_round_stats(h, rs.reshape(-1))
        snaps = []
        if snapshots:
            for r in range(R):
                c, l = lib.oracle_snapshot_clauses(h, r), lib.oracle_snapshot_literals(h, r)
                b = np.empty(c, np.uint32); sg = np.empty(c, np.uint32)
                of = np.empty(c + 1, np.uint64); li = np.empty(l, np.uint32)
                lib.oracle_copy_snapshot(h, r, b, sg, of, li)
                snaps.append(sgd.Dump.from_arrays(max_var, 2, b, sg, of, li, np.zeros(max_var + 1, np.uint8),
                                                  np.empty(0, np.uint32), np.empty(0, np.uint32)))
            elections = []
            for i in range(lib.oracle_num_elections(h)):
                e = np.empty(lib.oracle_election_size(h, i), np.uint32)
                lib.oracle_copy_election(h, i, e)
                elections.append(e)
            d.extra["elections"] = elections
        if proof:
            chunks = []
            for i in range(lib.oracle_proof_chunks(h)):
                b = np.empty(lib.oracle_proof_chunk_size(h, i), np.uint8)
                if len(b):
                    lib.oracle_copy_proof_chunk(h, i, b)
                chunks.append(b.tobytes())
            d.extra["proof"] = chunks
            d.extra["proof_cap"] = lib.oracle_proof_capacity(h)
    finally:
        lib.oracle_destroy(h)
    return d, rs, snaps


# ------------------------------------------------------------------ DRAT (binary) helpers for the proof-stream tests
def drat_parse(chunk: bytes):
    """binary DRAT -> [(b'a'|b'd', tuple of literals in the 2*var+sign encoding)]  (proof.cu:123-157 is the reader)"""
    out = []
    i, n = 0, len(chunk)
    while i < n:
        kind = chunk[i:i + 1]
        assert kind in (b"a", b"d"), f"bad line prefix {kind!r} at byte {i}"
        i += 1
        lits = []
        while True:
            assert i < n, "truncated proof line"
            if chunk[i] == 0:
                i += 1
                break
            v, shift = 0, 0
            while True:
                b = chunk[i]; i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not (b & 0x80):
                    break
            lits.append(v)
        assert lits, "empty proof clause"
        out.append((kind, tuple(lits)))
    return out


def drat_canonical(chunk: bytes):
    """order-free form of a chunk: sorted (kind, sorted literals) lines - the reference appends the lines of
    different variables in whatever order their threads reserve space (cuVecB::jump)."""
    return sorted((k, tuple(sorted(l))) for k, l in drat_parse(chunk))


class RupChecker:
    """Forward DRAT check, additions only by reverse unit propagation (every line the simplifier adds -
    strengthened clauses, substituted clauses, resolvents - is RUP).  Small formulas only: naive propagation."""

    def __init__(self, clauses):
        self.db = {}
        for c in clauses:
            self._add(tuple(sorted(set(c))))

    def _add(self, c):
        self.db[c] = self.db.get(c, 0) + 1

    def has(self, c):
        return self.db.get(tuple(sorted(set(c))), 0) > 0

    def delete(self, c):
        c = tuple(sorted(set(c)))
        if self.db.get(c, 0) <= 0:
            return False
        self.db[c] -= 1
        if not self.db[c]:
            del self.db[c]
        return True

    def rup(self, c):
        """True iff unit propagation on the database refutes the negation of clause c."""
        cs = set(c)
        if any((l ^ 1) in cs for l in cs):
            return True   # tautology
        assign = {l >> 1: (l & 1) for l in cs}   # every literal of c false: var = sign bit (2v+1 = negative literal)

        def val(l):
            a = assign.get(l >> 1)
            return None if a is None else (a == 1 - (l & 1))

        changed = True
        while changed:
            changed = False
            for cl in self.db:
                unassigned, n_un, sat = None, 0, False
                for l in cl:
                    v = val(l)
                    if v is True:
                        sat = True
                        break
                    if v is None:
                        n_un += 1
                        unassigned = l
                if sat:
                    continue
                if n_un == 0:
                    return True   # conflict
                if n_un == 1:
                    assign[unassigned >> 1] = 1 - (unassigned & 1)
                    changed = True
        return False

    def add(self, c):
        self._add(tuple(sorted(set(c))))
