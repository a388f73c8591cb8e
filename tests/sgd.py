"""Reader / canonicaliser for "SGD1" simplifier dumps (test infrastructure).

The same dump layout is written by
  * oracle/ref/ref_driver.cpp   -- the unmodified reference GPU solver (golden vectors),
  * oracle/sigma_oracle.cpp     -- the CPU restatement,
  * parafrost_b200 (sigma_store) -- the CUDA engine, through `Dump.from_arrays`.

Layout (uint32 little endian): 12-word header
  magic 'SGD1', maxVar, cnfstate, nClauses, nDataWords, nElim, nResolved, nTrail,
  numClauses, numLiterals, simpstate, pad
followed by the clause records {bits, sig, size, lits[size]} (the reference's SCLAUSE,
src/gpu/sclause.cuh:37-42), the eliminated bytes (word padded), the resolved stack
(src/gpu/model.cuh:29-53) and the root-level trail.
"""
from __future__ import annotations

import gzip
import json
from dataclasses import dataclass, field

import numpy as np

MAGIC = 0x31444753
U64 = np.uint64


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser, vectorised (wraps mod 2**64)."""
    x = x.astype(U64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> U64(30)
        x *= U64(0xBF58476D1CE4E5B9)
        x ^= x >> U64(27)
        x *= U64(0x94D049BB133111EB)
        x ^= x >> U64(31)
    return x


@dataclass
class Dump:
    max_var: int
    cnfstate: int
    num_clauses_inf: int
    num_literals_inf: int
    simpstate: int
    bits: np.ndarray      # per clause: st:2, f:1, a:1, u:2, lbd:26
    sig: np.ndarray
    size: np.ndarray
    offs: np.ndarray      # uint64 [C+1] into lits
    lits: np.ndarray
    eliminated: np.ndarray  # uint8 [maxVar+1]
    resolved: np.ndarray
    trail: np.ndarray
    extra: dict = field(default_factory=dict)

    # ------------------------------------------------------------------ io
    @staticmethod
    def load(path: str) -> "Dump":
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            raw = f.read()
        w = np.frombuffer(raw, dtype=np.uint32)
        assert int(w[0]) == MAGIC, "not an SGD1 dump"
        max_var, cnfstate, ncls, ndata, nelim, nres, ntrail, nc_inf, nl_inf, simpstate = (int(x) for x in w[1:11])
        data = w[12:12 + ndata]
        pos = 12 + ndata
        elim_words = (nelim + 3) // 4
        elim = np.frombuffer(w[pos:pos + elim_words].tobytes(), dtype=np.uint8)[:nelim].copy()
        pos += elim_words
        resolved = w[pos:pos + nres].copy()
        pos += nres
        trail = w[pos:pos + ntrail].copy()
        # walk the clause records
        bits = np.empty(ncls, np.uint32)
        sig = np.empty(ncls, np.uint32)
        size = np.empty(ncls, np.uint32)
        starts = np.empty(ncls, np.int64)
        p = 0
        dl = data.tolist() if ncls < 200000 else None
        if dl is not None:
            for i in range(ncls):
                bits[i] = dl[p]; sig[i] = dl[p + 1]; sz = dl[p + 2]; size[i] = sz
                starts[i] = p + 3
                p += 3 + sz
        else:  # vectorised walk by repeated doubling is overkill: plain loop over numpy scalars
            d = data
            for i in range(ncls):
                sz = int(d[p + 2])
                bits[i] = d[p]; sig[i] = d[p + 1]; size[i] = sz
                starts[i] = p + 3
                p += 3 + sz
        assert p == ndata, (p, ndata)
        offs = np.zeros(ncls + 1, np.uint64)
        np.cumsum(size, out=offs[1:])
        total = int(offs[-1])
        if ncls:
            idx = np.repeat(starts - offs[:-1].astype(np.int64), size.astype(np.int64)) + np.arange(total, dtype=np.int64)
            lits = data[idx].copy()
        else:
            lits = np.empty(0, np.uint32)
        return Dump(max_var, cnfstate, nc_inf, nl_inf, simpstate, bits, sig, size, offs, lits, elim, resolved, trail)

    @staticmethod
    def from_arrays(max_var, cnfstate, bits, sig, offs, lits, eliminated, resolved, trail,
                    num_clauses_inf=None, num_literals_inf=None, simpstate=0) -> "Dump":
        offs = np.asarray(offs, np.uint64)
        size = np.diff(offs).astype(np.uint32)
        return Dump(int(max_var), int(cnfstate),
                    int(len(size) if num_clauses_inf is None else num_clauses_inf),
                    int(len(lits) if num_literals_inf is None else num_literals_inf), int(simpstate),
                    np.asarray(bits, np.uint32), np.asarray(sig, np.uint32), size, offs,
                    np.asarray(lits, np.uint32), np.asarray(eliminated, np.uint8),
                    np.asarray(resolved, np.uint32), np.asarray(trail, np.uint32))

    def save(self, path: str) -> None:
        ncls = len(self.size)
        ndata = 3 * ncls + len(self.lits)
        data = np.empty(ndata, np.uint32)
        starts = 3 * np.arange(ncls, dtype=np.int64) + self.offs[:-1].astype(np.int64)
        data[starts] = self.bits
        data[starts + 1] = self.sig
        data[starts + 2] = self.size
        if len(self.lits):
            idx = np.repeat(starts + 3 - self.offs[:-1].astype(np.int64), self.size.astype(np.int64)) + np.arange(len(self.lits), dtype=np.int64)
            data[idx] = self.lits
        nelim = len(self.eliminated)
        elim = np.zeros(((nelim + 3) // 4) * 4, np.uint8)
        elim[:nelim] = self.eliminated
        hdr = np.array([MAGIC, self.max_var, self.cnfstate, ncls, ndata, nelim, len(self.resolved), len(self.trail),
                        self.num_clauses_inf, self.num_literals_inf, self.simpstate, 0], np.uint32)
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "wb") as f:
            f.write(hdr.tobytes()); f.write(data.tobytes()); f.write(elim.tobytes())
            f.write(self.resolved.astype(np.uint32).tobytes()); f.write(self.trail.astype(np.uint32).tobytes())

    # ------------------------------------------------------------------ canonical views
    def clause_hashes(self, with_flags: bool = False) -> np.ndarray:
        """One 64-bit hash per clause over its (sorted) literal set [+ flags word and sig]."""
        ncls = len(self.size)
        if ncls == 0:
            return np.empty(0, U64)
        lm = _mix64(self.lits.astype(U64) + U64(0x9E3779B97F4A7C15))
        # literals inside a clause are sorted ascending by the engine, but hash as a *set* anyway
        with np.errstate(over="ignore"):
            starts = self.offs[:-1].astype(np.int64)
            nonempty = self.size > 0
            acc = np.zeros(ncls, U64)
            if len(lm):
                red = np.add.reduceat(lm, np.minimum(starts, len(lm) - 1))
                acc[nonempty] = red[nonempty]
            acc = acc + self.size.astype(U64) * U64(0xD6E8FEB86659FD93)
            if with_flags:
                acc = acc + _mix64(self.bits.astype(U64) | (self.sig.astype(U64) << U64(32)))
        return _mix64(acc)

    def fingerprint(self) -> dict:
        """Order-free and ordered fingerprints used by the golden summaries."""
        def ms(h):
            with np.errstate(over="ignore"):
                s = int(np.add.reduce(h, dtype=U64)) if len(h) else 0
                x = int(np.bitwise_xor.reduce(h)) if len(h) else 0
            return f"{s:016x}{x:016x}"

        def ordered(h):
            with np.errstate(over="ignore"):
                k = _mix64(np.arange(len(h), dtype=U64) + U64(1))
                return f"{int(np.add.reduce(h * k, dtype=U64)) if len(h) else 0:016x}"

        hl = self.clause_hashes(False)
        hf = self.clause_hashes(True)
        elim_vars = np.nonzero(self.eliminated & 1)[0].astype(U64)
        forced_vars = np.nonzero(self.eliminated & 4)[0].astype(U64)
        gh = self.group_hashes()   # == _mix64(hash_words(g)) for g in resolved_groups(), without the Python loops
        # the same stack record by record: with BCE the blocked-clause records (no closing unit) sit in front of whichever
        # group the atomic append put behind them, so runs with BCE compare this multiset instead of the groups
        rh = self.record_hashes()
        return {
            "max_var": self.max_var,
            "cnfstate": self.cnfstate,
            "clauses": int(len(self.size)),
            "literals": int(len(self.lits)),
            "num_clauses_inf": self.num_clauses_inf,
            "num_literals_inf": self.num_literals_inf,
            "eliminated": int(len(elim_vars)),
            "forced": int(len(forced_vars)),
            "resolved_words": int(len(self.resolved)),
            "resolved_groups": int(len(gh)),
            "trail": int(len(self.trail)),
            "h_lits_multiset": ms(hl),
            "h_full_multiset": ms(hf),
            "h_lits_ordered": ordered(hl),
            "h_full_ordered": ordered(hf),
            "h_eliminated": ms(_mix64(elim_vars)),
            "h_forced": ms(_mix64(forced_vars)),
            "h_resolved_groups": ms(gh),
            "h_resolved_records": ms(rh),
            "h_trail_multiset": ms(_mix64(self.trail.astype(U64))),
        }

    def canonical_clauses(self) -> list[tuple[int, ...]]:
        """Sorted list of literal tuples (small dumps only)."""
        o = self.offs.astype(np.int64)
        l = self.lits.tolist()
        return sorted(tuple(l[o[i]:o[i + 1]]) for i in range(len(self.size)))

    def ordered_clauses(self) -> list[tuple[int, ...]]:
        o = self.offs.astype(np.int64)
        l = self.lits.tolist()
        return [tuple(l[o[i]:o[i + 1]]) for i in range(len(self.size))]

    def eliminated_vars(self) -> list[int]:
        return np.nonzero(self.eliminated & 1)[0].tolist()

    def record_ends(self) -> np.ndarray:
        """ends[k] = one past record k of the witness stack.  Records are `[lits..., size]`: the sizes sit at the END, so
        the boundaries are found walking back from the top - in C when the oracle library is there (a 100 M-word stack takes
        a minute of Python otherwise)."""
        n = len(self.resolved)
        if not n:
            return np.empty(0, np.int64)
        r = np.ascontiguousarray(self.resolved, np.uint32)
        try:
            import helpers
            lib = helpers.oracle_lib()
            cnt = lib.oracle_record_ends(r.ctypes.data, n, None)
            assert cnt != 0xFFFFFFFFFFFFFFFF, "corrupt resolved stack"
            ends = np.empty(cnt, np.uint64)
            lib.oracle_record_ends(r.ctypes.data, n, ends.ctypes.data)
            return ends.astype(np.int64)
        except (ImportError, OSError, AttributeError):
            pass
        rl = r.tolist()
        ends = []
        p = n
        while p > 0:
            sz = rl[p - 1]
            assert 0 < sz < p + 1, "corrupt resolved stack"
            ends.append(p)
            p -= 1 + sz
        return np.array(ends[::-1], np.int64)

    def record_hashes(self) -> np.ndarray:
        """One 64-bit hash per record `[lits..., size]` of the witness stack (position inside the record matters: the
        witness literal comes first)."""
        n = len(self.resolved)
        if not n:
            return np.empty(0, U64)
        ends = self.record_ends()
        starts = np.concatenate((np.zeros(1, np.int64), ends[:-1]))
        pos = np.arange(n, dtype=np.int64) - np.repeat(starts, ends - starts)
        with np.errstate(over="ignore"):
            w = _mix64(self.resolved.astype(U64) + (pos.astype(U64) << U64(32)))
            return _mix64(np.add.reduceat(w, starts))

    def group_hashes(self) -> np.ndarray:
        """_mix64(hash_words(g)) for every group g of resolved_groups(): a group is the run of records up to and including a
        unit record - a contiguous slice of the stack -, records after the last unit are groups of their own."""
        n = len(self.resolved)
        if not n:
            return np.empty(0, U64)
        r = np.ascontiguousarray(self.resolved, np.uint32)
        ends = self.record_ends()
        rstarts = np.concatenate((np.zeros(1, np.int64), ends[:-1]))
        unit = r[ends - 1] == 1
        gends = ends[unit]
        gstarts = np.concatenate((np.zeros(1, np.int64), gends[:-1])) if len(gends) else np.empty(0, np.int64)
        last = int(gends[-1]) if len(gends) else 0
        loose = rstarts >= last            # records behind the last unit (blocked clauses of the last round)
        gstarts = np.concatenate((gstarts, rstarts[loose])).astype(np.uint64)
        gends = np.concatenate((gends, ends[loose])).astype(np.uint64)
        out = np.empty(len(gends), U64)
        try:
            import helpers
            helpers.oracle_lib().oracle_hash_segments(r.ctypes.data, gstarts.ctypes.data, gends.ctypes.data, len(gends), out.ctypes.data)
        except (ImportError, OSError, AttributeError):
            rl = r.tolist()
            for k in range(len(gends)):
                out[k] = hash_words(rl[int(gstarts[k]):int(gends[k])])
        return _mix64(out)

    def resolved_groups(self) -> list[tuple[int, ...]]:
        """Split the witness stack into per-variable groups.

        Records are `[lits..., size]`; a group is a run of clause records closed by the
        witness unit `[lit, 1]`.  Groups from one round may interleave in any order in the
        reference (atomic jump, src/gpu/vector.cu:85-90), so they are compared as a multiset:
        each group is returned as a flat tuple (witness-first clause records kept in order).
        """
        r = self.resolved.tolist()
        recs = []
        p = len(r)
        while p > 0:
            sz = r[p - 1]
            assert 0 < sz < p + 1, "corrupt resolved stack"
            recs.append(tuple(r[p - 1 - sz:p - 1]))
            p -= 1 + sz
        recs.reverse()
        groups, cur = [], []
        for rec in recs:
            cur.append(rec)
            if len(rec) == 1:
                groups.append(tuple(x for rr in cur for x in (*rr, len(rr))))
                cur = []
        if cur:  # blocked-clause records (BCE) are not closed by a unit: one group each
            for rec in cur:
                groups.append(tuple((*rec, len(rec))))
        return sorted(groups)


    def canonical_witness(self):
        """Order-independent form of the witness stack for runs with BCE.  A BVE group is `k` clause
        records whose FIRST literal is the witness literal p (elimination.cuh:505-550) closed by the
        unit record [FLIP(p), 1]; a blocked-clause record (blocked.cuh) has no closing unit and - the
        append order inside a round being arbitrary (atomic jump, vector.cu:85-90) - may sit in front
        of any group of the next round.  Returns (sorted groups, sorted loose records): a unit claims
        only the directly preceding records that start with its flipped literal."""
        r = self.resolved.tolist()
        recs = []
        p = len(r)
        while p > 0:
            sz = r[p - 1]
            assert 0 < sz < p + 1, "corrupt resolved stack"
            recs.append(tuple(r[p - 1 - sz:p - 1]))
            p -= 1 + sz
        recs.reverse()
        claimed = [False] * len(recs)
        groups = []
        for t, rec in enumerate(recs):
            if len(rec) != 1:
                continue
            claimed[t] = True
            want = rec[0] ^ 1
            k = t - 1
            mine = []
            while k >= 0 and not claimed[k] and len(recs[k]) > 1 and recs[k][0] == want:
                claimed[k] = True
                mine.append(recs[k])
                k -= 1
            groups.append((rec[0], tuple(sorted(mine))))
        loose = sorted(recs[t] for t in range(len(recs)) if not claimed[t])
        return sorted(groups), loose


def hash_words(words) -> int:
    h = 0xCBF29CE484222325
    for w in words:
        h ^= int(w) & 0xFFFFFFFF
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def compare(a: Dump, b: Dump, ordered: bool = True, flags: bool = True) -> list[str]:
    """Return a list of human-readable mismatches between two dumps (empty = parity)."""
    if a.cnfstate == 0 and b.cnfstate == 0 and a.max_var == b.max_var:
        return []   # UNSAT by propagation (elimbcp.cu:178): the caller learns the empty clause; which variables the parallel
                    # BFS had assigned when the conflict surfaced depends on its schedule (in the reference too) and is unused
    fa, fb = a.fingerprint(), b.fingerprint()
    keys = ["max_var", "cnfstate", "clauses", "literals", "eliminated", "forced", "resolved_words",
            "resolved_groups", "trail", "h_lits_multiset", "h_eliminated", "h_forced",
            "h_resolved_groups", "h_trail_multiset"]
    if flags:
        keys.append("h_full_multiset")
    if ordered:
        keys.append("h_lits_ordered")
        if flags:
            keys.append("h_full_ordered")
    diff = [f"{k}: {fa[k]} != {fb[k]}" for k in keys if fa[k] != fb[k]]
    if len(diff) == 1 and diff[0].startswith("h_resolved_groups") and a.canonical_witness() == b.canonical_witness():
        return []   # same records, blocked-clause records attached to different neighbours (see canonical_witness)
    return diff


if __name__ == "__main__":
    import sys
    out = {}
    for p in sys.argv[1:]:
        out[p] = Dump.load(p).fingerprint()
    print(json.dumps(out, indent=1, sort_keys=True))
