"""ctypes binding of libsigma_b200.so (include/sigma.h).

Mirrors the simplifier surface of the reference's ``class Solver`` (src/gpu/solver.hpp:674-789):
``optSimp`` (options), ``awaken`` (load), ``simplify`` (the round loop), ``cacheCNF`` /
``cacheResolved`` / ``cacheEliminated`` (store), ``freeSimp``.  Flags carry the reference's CLI
names (``-no-ere``, ``--phases=K`` ...).  There is no fallback path: if the CUDA library cannot
be loaded this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import lib_path

UNSAT, SAT, UNSOLVED = 0, 1, 2


class SigmaOpts(C.Structure):
    _fields_ = [
        ("phases", C.c_int32), ("ve_en", C.c_int32), ("ve_plus_en", C.c_int32), ("sub_en", C.c_int32),
        ("bce_en", C.c_int32), ("ere_en", C.c_int32), ("all_en", C.c_int32),
        ("mu_pos", C.c_uint32), ("mu_neg", C.c_uint32), ("lcve_min_vars", C.c_uint32),
        ("lcve_max_occurs", C.c_uint32), ("lcve_clause_max", C.c_int32), ("phase_lits_min", C.c_int32),
        ("shrink_rate", C.c_int32), ("lits_mul", C.c_double),
        ("ve_fun_en", C.c_int32), ("ve_lbound_en", C.c_int32), ("ve_clause_max", C.c_uint32),
        ("xor_max_arity", C.c_uint32), ("ere_clause_max", C.c_int32), ("ere_max_occurs", C.c_uint32),
        ("sub_max_occurs", C.c_uint32), ("bce_max_occurs", C.c_uint32), ("sh_max_bve_out1", C.c_uint32),
        ("sigma_calls", C.c_int32), ("final_gc", C.c_int32), ("profile", C.c_int32), ("aggr_cnf_sort", C.c_int32),
        ("proof_en", C.c_int32), ("lcve_fast", C.c_int32), ("log_reductions", C.c_int32),
    ]


class RoundReport(C.Structure):
    _fields_ = [
        ("round", C.c_uint32), ("kind", C.c_uint32), ("elected", C.c_uint32), ("eliminated", C.c_uint32),
        ("resolvents", C.c_uint32), ("units", C.c_uint32), ("propagated", C.c_uint32), ("gc", C.c_uint32),
        ("clauses", C.c_uint64), ("literals", C.c_uint64), ("literals_in", C.c_uint64),
        ("ms", C.c_float), ("trail_added", C.c_uint32),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Report(C.Structure):
    _fields_ = [
        ("cnfstate", C.c_int32), ("simpstate", C.c_int32), ("rounds", C.c_uint32), ("eliminated_vars", C.c_uint32),
        ("clauses", C.c_uint64), ("literals", C.c_uint64), ("clauses_in", C.c_uint64), ("literals_in", C.c_uint64),
        ("resolved_words", C.c_uint64), ("trail_units", C.c_uint64), ("ms_total", C.c_double),
        ("stage_ms", C.c_float * 16), ("kernel_launches", C.c_uint64), ("ms_device", C.c_double),
    ]

    STAGES = ["vo", "sig", "io", "gc", "cot", "sot", "rot", "ve", "sub", "bce", "ere", "prop", "lcve", "cnt"]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "stage_ms"}
        d["stage_ms"] = {s: float(self.stage_ms[i]) for i, s in enumerate(self.STAGES)}
        return d


class StageReduction(C.Structure):
    _fields_ = [("round", C.c_uint32), ("stage", C.c_uint32), ("vars_removed", C.c_uint32), ("pad", C.c_uint32),
                ("clauses_before", C.c_uint64), ("literals_before", C.c_uint64), ("clauses", C.c_uint64), ("literals", C.c_uint64)]
    STAGES = ["BCP", "SUB", "BVE", "BCE", "ERE"]


class DeviceCnf(C.Structure):
    """sigma_device_cnf: device pointers into the context's arena (Solver::getDeviceCNF / getVars, solver.hpp:694-705)."""
    _fields_ = [
        ("device", C.c_int32), ("stream", C.c_void_p), ("max_var", C.c_uint32), ("clause_slots", C.c_uint32), ("pool_words", C.c_uint64),
        ("live_clauses", C.c_uint64), ("live_literals", C.c_uint64), ("headers", C.c_void_p), ("literals", C.c_void_p),
        ("ot_start", C.c_void_p), ("ot_size", C.c_void_p), ("ot_entries", C.c_void_p), ("eliminated", C.c_void_p), ("vstate", C.c_void_p),
        ("vorg", C.c_void_p), ("elected", C.c_void_p), ("num_elected", C.c_uint32), ("units", C.c_void_p), ("resolved", C.c_void_p),
        ("resolved_words", C.c_uint64), ("trail", C.c_void_p), ("trail_units", C.c_uint64),
    ]


class SigmaError(RuntimeError):
    pass


_lib = None
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

# every symbol include/sigma.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sigma_default_opts", "sigma_normalize_opts", "sigma_create", "sigma_destroy", "sigma_set_opts", "sigma_set_stream", "sigma_load", "sigma_load32", "sigma_load_sclauses",
    "sigma_run", "sigma_begin", "sigma_round", "sigma_finish", "sigma_num_rounds", "sigma_round_reports",
    "sigma_result_sizes", "sigma_store", "sigma_store_compact", "sigma_store_sclauses", "sigma_snapshot", "sigma_debug_elected",
    "sigma_set_proof_sink", "sigma_proof_chunks", "sigma_proof_chunk_size", "sigma_proof_chunk_copy",
    "sigma_debug_hist", "sigma_kernel_profile", "sigma_kernel_times", "sigma_kernel_stats", "sigma_trail_info", "sigma_copy_trail", "sigma_pinned_alloc", "sigma_pinned_free",
    "sigma_device_view", "sigma_continue", "sigma_reduction_log", "sigma_memory", "sigma_last_error", "sigma_version", "sigma_stage_prep", "sigma_stage_histogram",
]


def lib():
    """Loads the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise SigmaError(f"{path} is missing: run `python -m parafrost_b200.build` (nvcc, sm_100a); there is no CPU fallback")
        L = C.CDLL(path)
        P = C.c_void_p
        L.sigma_default_opts.argtypes = [C.POINTER(SigmaOpts)]
        L.sigma_normalize_opts.argtypes = [C.POINTER(SigmaOpts)]
        L.sigma_create.argtypes = [C.c_int, C.POINTER(SigmaOpts), C.POINTER(P)]
        L.sigma_destroy.argtypes = [P]
        L.sigma_set_opts.argtypes = [P, C.POINTER(SigmaOpts)]
        L.sigma_set_stream.argtypes = [P, P]
        L.sigma_load.argtypes = [P, C.c_uint32, C.c_uint64, P, P, P, P, P, P]
        L.sigma_load32.argtypes = [P, C.c_uint32, C.c_uint64, P, P, P, P, P, P]
        L.sigma_load_sclauses.argtypes = [P, C.c_uint32, C.c_uint64, P, C.c_uint64, P, P, P, P]
        L.sigma_run.argtypes = [P, C.POINTER(Report)]
        L.sigma_begin.argtypes = [P]
        L.sigma_round.argtypes = [P, C.POINTER(RoundReport), C.POINTER(C.c_int)]
        L.sigma_finish.argtypes = [P, C.POINTER(Report)]
        L.sigma_num_rounds.argtypes = [P]; L.sigma_num_rounds.restype = C.c_uint32
        L.sigma_round_reports.argtypes = [P, C.POINTER(RoundReport), C.c_uint32]
        L.sigma_result_sizes.argtypes = [P] + [C.POINTER(C.c_uint64)] * 4
        L.sigma_store.argtypes = [P, P, P, P, P, P, P, P]
        L.sigma_store_compact.argtypes = [P, P, P, P, P, P, P]
        L.sigma_store_sclauses.argtypes = [P, P, P]
        L.sigma_snapshot.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.sigma_debug_elected.argtypes = [P, P, C.POINTER(C.c_uint32)]
        L.sigma_debug_hist.argtypes = [P, P]
        L.sigma_set_proof_sink.argtypes = [P, P, P]
        L.sigma_proof_chunks.argtypes = [P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.sigma_proof_chunk_size.argtypes = [P, C.c_uint32, C.POINTER(C.c_uint64)]
        L.sigma_proof_chunk_copy.argtypes = [P, C.c_uint32, P]
        L.sigma_kernel_profile.argtypes = [P, C.c_int]
        L.sigma_kernel_times.argtypes = [P, P, P, P, C.POINTER(C.c_uint32)]
        L.sigma_trail_info.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.sigma_copy_trail.argtypes = [P, C.c_uint64, C.c_uint64, P]
        L.sigma_pinned_alloc.argtypes = [C.c_size_t]; L.sigma_pinned_alloc.restype = C.c_void_p
        L.sigma_pinned_free.argtypes = [C.c_void_p]
        L.sigma_device_view.argtypes = [P, C.POINTER(DeviceCnf)]
        L.sigma_reduction_log.argtypes = [P, P, C.POINTER(C.c_uint32)]
        L.sigma_continue.argtypes = [P, C.c_uint64, P, P, P, P, P]
        L.sigma_kernel_stats.argtypes = [P, P, P, P, P, C.POINTER(C.c_uint32)]
        L.sigma_memory.argtypes = [P] + [C.POINTER(C.c_uint64)] * 3
        L.sigma_last_error.argtypes = [P]; L.sigma_last_error.restype = C.c_char_p
        L.sigma_version.restype = C.c_char_p
        L.sigma_stage_prep.argtypes = [C.c_int, C.c_uint64, _u32p, _u64p, _u32p]
        L.sigma_stage_histogram.argtypes = [C.c_int, C.c_uint64, _u32p, C.c_uint32, _u32p]
        _lib = L
    return _lib


FLAG_MAP = {  # the reference's CLI flags (src/gpu/options.cpp:24-43, options.cu:36-60)
    "-no-ere": {"ere_en": 0}, "-ere": {"ere_en": 1}, "-no-vefunction": {"ve_fun_en": 0}, "-bce": {"bce_en": 1},
    "-all": {"all_en": 1}, "-no-sub": {"sub_en": 0}, "-no-veextend": {"ve_plus_en": 0}, "-no-ve": {"ve_en": 0},
    "-velitsbound": {"ve_lbound_en": 1}, "-profilegpu": {"profile": 1}, "-aggresivesort": {"aggr_cnf_sort": 1}, "-no-lcvefast": {"lcve_fast": 0}, "-lcvefast": {"lcve_fast": 1}, "-quiet": {},
    "-proof": {"proof_en": 1},
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
            if k == "--verbose":                       # LOGREDALL / LOGREDCL tables need --verbose >= 2 (logging.hpp:152-158)
                if int(v) >= 2:
                    o["log_reductions"] = 1
            elif k in VALUE_FLAGS:
                o[VALUE_FLAGS[k]] = float(v) if k == "--literalsmul" else int(v)
            elif k not in ("--mapperc", "--ereminthreads"):
                raise ValueError(f"unknown flag {f}")
        else:
            o.update(FLAG_MAP[f])
    return o


def make_opts(**over) -> SigmaOpts:
    o = SigmaOpts()
    lib().sigma_default_opts(C.byref(o))
    for k, v in over.items():
        setattr(o, k, v)
    lib().sigma_normalize_opts(C.byref(o))
    return o


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Simplifier:
    """One SIGmA context on one device (the simplifier half of the reference's ``Solver``)."""

    def __init__(self, device: int = 0, flags=(), **opts):
        self._h = C.c_void_p()
        self._lib = lib()
        o = dict(opts_from_flags(flags))
        o.update(opts)
        self.opts = make_opts(**o)
        rc = self._lib.sigma_create(device, C.byref(self.opts), C.byref(self._h))
        if rc:
            raise SigmaError(f"sigma_create failed ({rc}): no usable CUDA device {device}?")
        self.max_var = 0

    # Solver::optSimp
    def optSimp(self, flags=(), **opts):
        o = dict(opts_from_flags(flags))
        o.update(opts)
        self.opts = make_opts(**o)
        self._check(self._lib.sigma_set_opts(self._h, C.byref(self.opts)))

    def set_stream(self, cuda_stream: int):
        """Run this context on the caller's CUDA stream (e.g. ``torch.cuda.Stream().cuda_stream``)."""
        self._check(self._lib.sigma_set_stream(self._h, C.c_void_p(cuda_stream)))

    def _check(self, rc):
        if rc:
            raise SigmaError(f"sigma error {rc}: {self._lib.sigma_last_error(self._h).decode()}")

    # Solver::awaken (host half)
    def load(self, max_var, lits, offs, meta=None, vorg=None, vstate=None, assumed=None):
        """CSR clause list; `offs` of dtype uint32 goes through sigma_load32 (half the offset bytes over PCIe)."""
        o32 = isinstance(offs, np.ndarray) and offs.dtype == np.uint32
        self._keep = [np.ascontiguousarray(lits, np.uint32), np.ascontiguousarray(offs, np.uint32 if o32 else np.uint64),
                      None if meta is None else np.ascontiguousarray(meta, np.uint32),
                      None if vorg is None else np.ascontiguousarray(vorg, np.uint32),
                      None if vstate is None else np.ascontiguousarray(vstate, np.uint8),
                      None if assumed is None else np.ascontiguousarray(assumed, np.uint8)]
        k = self._keep
        self.max_var = int(max_var)
        self._check((self._lib.sigma_load32 if o32 else self._lib.sigma_load)(self._h, self.max_var, len(k[1]) - 1, *[_ptr(a) for a in k]))

    def load_sclauses(self, max_var, data_words, refs, vorg=None, vstate=None, assumed=None):
        """The reference's host mirror (SCLAUSE record stream + uint64 refs, cnf.cuh:82-97), as produced
        by CNF::newClause on the host or by store_sclauses()."""
        self._keep = [np.ascontiguousarray(data_words, np.uint32), np.ascontiguousarray(refs, np.uint64),
                      None if vorg is None else np.ascontiguousarray(vorg, np.uint32),
                      None if vstate is None else np.ascontiguousarray(vstate, np.uint8),
                      None if assumed is None else np.ascontiguousarray(assumed, np.uint8)]
        k = self._keep
        self.max_var = int(max_var)
        self._check(self._lib.sigma_load_sclauses(self._h, self.max_var, len(k[1]), _ptr(k[0]), len(k[0]), _ptr(k[1]),
                                                  _ptr(k[2]), _ptr(k[3]), _ptr(k[4])))

    def load_pointers(self, max_var, num_clauses, lits_ptr, offs_ptr, offs32=False):
        """Raw host pointers (e.g. pinned torch tensors) - used by bench.py's e2e leg."""
        self.max_var = int(max_var)
        fn = self._lib.sigma_load32 if offs32 else self._lib.sigma_load
        self._check(fn(self._h, self.max_var, num_clauses, lits_ptr, offs_ptr, None, None, None, None))

    # Solver::simplify
    def simplify(self) -> dict:
        rep = Report()
        self._check(self._lib.sigma_run(self._h, C.byref(rep)))
        return rep.asdict()

    def begin(self):
        self._check(self._lib.sigma_begin(self._h))

    def round(self):
        r = RoundReport()
        done = C.c_int(0)
        self._check(self._lib.sigma_round(self._h, C.byref(r), C.byref(done)))
        return r.asdict(), bool(done.value)

    def finish(self) -> dict:
        rep = Report()
        self._check(self._lib.sigma_finish(self._h, C.byref(rep)))
        return rep.asdict()

    def rounds(self):
        n = self._lib.sigma_num_rounds(self._h)
        arr = (RoundReport * max(n, 1))()
        self._check(self._lib.sigma_round_reports(self._h, arr, n))
        return [arr[i].asdict() for i in range(n)]

    # Solver::cacheCNF / cacheResolved / cacheEliminated
    def store(self, into: dict | None = None) -> dict:
        """cacheCNF + cacheResolved + cacheEliminated.  `into` may hold preallocated (e.g. pinned)
        arrays under the keys below, each at least as large as the result; views are returned."""
        nc, nl, nr, nt = (C.c_uint64() for _ in range(4))
        self._check(self._lib.sigma_result_sizes(self._h, C.byref(nc), C.byref(nl), C.byref(nr), C.byref(nt)))
        if into is not None:
            need = {"bits": nc.value, "sig": nc.value, "offs": nc.value + 1, "lits": nl.value,
                    "eliminated": self.max_var + 1, "resolved": nr.value, "trail": nt.value}
            for k, n in need.items():
                if len(into[k]) < n:
                    raise SigmaError(f"store: buffer {k} holds {len(into[k])} < {n}")
            self._check(self._lib.sigma_store(self._h, *[_ptr(into[k]) for k in ("bits", "sig", "offs", "lits", "eliminated", "resolved", "trail")]))
            return {k: into[k][:n] for k, n in need.items()}
        out = {
            "bits": np.empty(nc.value, np.uint32), "sig": np.empty(nc.value, np.uint32),
            "offs": np.zeros(nc.value + 1, np.uint64), "lits": np.empty(nl.value, np.uint32),
            "eliminated": np.zeros(self.max_var + 1, np.uint8), "resolved": np.empty(nr.value, np.uint32),
            "trail": np.empty(nt.value, np.uint32),
        }
        self._check(self._lib.sigma_store(self._h, *[_ptr(out[k]) for k in ("bits", "sig", "offs", "lits", "eliminated", "resolved", "trail")]))
        return out

    def store_compact(self, into: dict | None = None) -> dict:
        """What writeBackCNF -> newClause(SCLAUSE&) reads (sclause.cpp:22-55): word 0, size and literals of every clause -
        8 + 4|c| bytes per clause over PCIe instead of 16 + 4|c| (no signatures, no 64-bit offsets)."""
        nc, nl, nr, nt = (C.c_uint64() for _ in range(4))
        self._check(self._lib.sigma_result_sizes(self._h, C.byref(nc), C.byref(nl), C.byref(nr), C.byref(nt)))
        need = {"bits": nc.value, "sizes": nc.value, "lits": nl.value, "eliminated": self.max_var + 1, "resolved": nr.value, "trail": nt.value}
        if into is None:
            into = {k: (np.zeros(n, np.uint8) if k == "eliminated" else np.empty(n, np.uint32)) for k, n in need.items()}
        for k, n in need.items():
            if len(into[k]) < n:
                raise SigmaError(f"store_compact: buffer {k} holds {len(into[k])} < {n}")
        self._check(self._lib.sigma_store_compact(self._h, *[_ptr(into[k]) for k in ("bits", "sizes", "lits", "eliminated", "resolved", "trail")]))
        return {k: into[k][:n] for k, n in need.items()}

    def device_view(self) -> DeviceCnf:
        """Device pointers of the resident result (simplify(skip_transfer_to_host))."""
        v = DeviceCnf()
        self._check(self._lib.sigma_device_view(self._h, C.byref(v)))
        return v

    def continue_resident(self, new_lits=None, new_offs=None, new_meta=None, vstate=None, assumed=None):
        """The next inprocessing call on the resident result (sigma_continue): only the clauses added since travel."""
        k = [None if a is None else np.ascontiguousarray(a, dt) for a, dt in
             ((new_lits, np.uint32), (new_offs, np.uint64), (new_meta, np.uint32), (vstate, np.uint8), (assumed, np.uint8))]
        n = 0 if k[1] is None else len(k[1]) - 1
        self._keep2 = k
        self._check(self._lib.sigma_continue(self._h, n, _ptr(k[0]), _ptr(k[1]), _ptr(k[2]), _ptr(k[3]), _ptr(k[4])))

    def reduction_log(self):
        """LOGREDALL / LOGREDCL tables of the last run (needs log_reductions=1): [{round, stage, vars_removed, clauses, literals, ...}]"""
        n = C.c_uint32(0)
        self._check(self._lib.sigma_reduction_log(self._h, None, C.byref(n)))
        arr = (StageReduction * max(n.value, 1))()
        m = C.c_uint32(n.value)
        self._check(self._lib.sigma_reduction_log(self._h, arr, C.byref(m)))
        return [{"round": e.round, "stage": StageReduction.STAGES[e.stage], "vars_removed": e.vars_removed, "clauses_before": e.clauses_before,
                 "literals_before": e.literals_before, "clauses": e.clauses, "literals": e.literals} for e in arr[: n.value]]

    def trail_info(self):
        tot, frm, cnt, seeds = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib.sigma_trail_info(self._h, C.byref(tot), C.byref(frm), C.byref(cnt), C.byref(seeds)))
        return {"total": tot.value, "last_from": frm.value, "last_count": cnt.value, "last_seeds": seeds.value}

    def store_sclauses(self):
        nc, nl, nr, nt = (C.c_uint64() for _ in range(4))
        self._check(self._lib.sigma_result_sizes(self._h, C.byref(nc), C.byref(nl), C.byref(nr), C.byref(nt)))
        data = np.empty(3 * nc.value + nl.value, np.uint32)
        refs = np.empty(nc.value, np.uint64)
        self._check(self._lib.sigma_store_sclauses(self._h, _ptr(data), _ptr(refs)))
        return data, refs

    def snapshot(self) -> dict:
        """Live clause list right now (per-round parity checks)."""
        return self.store()

    # cuPROOF::cacheProof / writeProof (proof.cu:160-199, 232-247)
    PROOF_SINK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_uint8), C.c_uint64)

    def set_proof_sink(self, fn):
        """fn(bytes) is called with every round's chunk of the device DRAT stream (flag ``-proof``)."""
        if fn is None:
            self._sink = None
            self._check(self._lib.sigma_set_proof_sink(self._h, None, None))
            return
        self._sink = self.PROOF_SINK(lambda _u, p, n: fn(C.string_at(p, n)))
        self._check(self._lib.sigma_set_proof_sink(self._h, C.cast(self._sink, C.c_void_p), None))

    def proof_chunks(self):
        """-> ([one bytes object per cacheProof/writeProof point of the last run], capacity in bytes)"""
        n, tot, cap = C.c_uint32(), C.c_uint64(), C.c_uint32()
        self._check(self._lib.sigma_proof_chunks(self._h, C.byref(n), C.byref(tot), C.byref(cap)))
        out = []
        for i in range(n.value):
            sz = C.c_uint64()
            self._check(self._lib.sigma_proof_chunk_size(self._h, i, C.byref(sz)))
            buf = np.empty(max(sz.value, 1), np.uint8)
            if sz.value:
                self._check(self._lib.sigma_proof_chunk_copy(self._h, i, _ptr(buf)))
            out.append(buf[: sz.value].tobytes())
        return out, cap.value

    def debug_elected(self):
        n = C.c_uint32(0)
        buf = np.zeros(self.max_var + 1, np.uint32)
        self._check(self._lib.sigma_debug_elected(self._h, _ptr(buf), C.byref(n)))
        return buf[: n.value].copy()

    def debug_hist(self):
        buf = np.zeros(2 * (self.max_var + 1), np.uint32)
        self._check(self._lib.sigma_debug_hist(self._h, _ptr(buf)))
        return buf

    def kernel_profile(self, enable: int = 1):
        """CUDA-event pair around every kernel launch, on the engine's own stream."""
        self._check(self._lib.sigma_kernel_profile(self._h, int(enable)))

    def kernel_times(self) -> dict:
        """-> {kernel name: (total ms, launches)} since kernel_profile(1)."""
        cap = 128
        names = C.create_string_buffer(64 * cap)
        ms = (C.c_float * cap)()
        cnt = (C.c_uint32 * cap)()
        n = C.c_uint32(cap)
        self._check(self._lib.sigma_kernel_times(self._h, names, ms, cnt, C.byref(n)))
        return {names.raw[64 * i:64 * (i + 1)].split(b"\0", 1)[0].decode(): (float(ms[i]), int(cnt[i])) for i in range(n.value)}

    def kernel_stats(self) -> dict:
        """-> {kernel name: (total ms, launches, algorithmic bytes)} since kernel_profile(1)."""
        cap = 128
        names = C.create_string_buffer(64 * cap)
        ms = (C.c_float * cap)()
        cnt = (C.c_uint32 * cap)()
        by = (C.c_double * cap)()
        n = C.c_uint32(cap)
        self._check(self._lib.sigma_kernel_stats(self._h, names, ms, cnt, by, C.byref(n)))
        return {names.raw[64 * i:64 * (i + 1)].split(b"\0", 1)[0].decode(): (float(ms[i]), int(cnt[i]), float(by[i])) for i in range(n.value)}

    def memory(self) -> dict:
        a, p, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._lib.sigma_memory(self._h, C.byref(a), C.byref(p), C.byref(m))
        return {"arena_bytes": a.value, "peak_used": p.value, "cuda_mallocs": m.value}

    # Solver::freeSimp
    def freeSimp(self):
        if self._h:
            self._lib.sigma_destroy(self._h)
            self._h = C.c_void_p()

    close = freeSimp

    def __del__(self):
        try:
            self.freeSimp()
        except Exception:
            pass


def stage_prep(lits, offs, device: int = 0):
    """prep_cnf_k through the C ABI: returns (sorted lits, sig)."""
    lits = np.ascontiguousarray(lits, np.uint32).copy()
    offs = np.ascontiguousarray(offs, np.uint64)
    sig = np.zeros(len(offs) - 1, np.uint32)
    rc = lib().sigma_stage_prep(device, len(offs) - 1, lits, offs, sig)
    if rc:
        raise SigmaError(f"sigma_stage_prep failed ({rc})")
    return lits, sig


def stage_histogram(lits, nbins, device: int = 0):
    lits = np.ascontiguousarray(lits, np.uint32)
    hist = np.zeros(nbins, np.uint32)
    rc = lib().sigma_stage_histogram(device, len(lits), lits, nbins, hist)
    if rc:
        raise SigmaError(f"sigma_stage_histogram failed ({rc})")
    return hist
