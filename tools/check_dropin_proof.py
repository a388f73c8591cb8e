#!/usr/bin/env python
"""End-to-end DRAT check of the drop-in (GPU box): the relinked `parafrost` CLI (oracle/_ref/parafrost_sigma = the reference's
host objects + integration/sigma_shim.cpp + libsigma_b200.so) solves a small UNSAT formula with -proof; the proof file - the
reference's CDCL lines, the units the shim replays per device prop(), and the engine's device stream in between - is checked
forward by reverse unit propagation (tests/helpers.RupChecker) and must end in the empty clause.
    python tools/check_dropin_proof.py [family seed args...]      -> one JSON line"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import helpers  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "parafrost_sigma")


def parse(raw):
    out, i, n = [], 0, len(raw)
    while i < n:
        kind = raw[i:i + 1]
        if kind not in (b"a", b"d"):
            raise ValueError(f"bad line prefix {kind!r} at byte {i}")
        i += 1
        lits = []
        while raw[i] != 0:
            v, shift = 0, 0
            while True:
                b = raw[i]; i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not (b & 0x80):
                    break
            lits.append(v)
        i += 1
        out.append((kind, tuple(lits)))
    return out


def check(fam, seed, args, flags=()):
    tmp = tempfile.mkdtemp()
    cnf, prf = os.path.join(tmp, "f.cnf"), os.path.join(tmp, "f.drat")
    V, lits, offs = helpers.gen_cnf(fam, seed, args, dimacs_path=cnf)
    r = subprocess.run([BIN, cnf, "-proof", "--proofout=" + prf] + list(flags), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
    ans = [l for l in out.splitlines() if l.startswith("s ")]
    res = {"instance": f"{fam}{tuple(args)} seed {seed}", "flags": list(flags), "answer": ans[-1][2:].strip() if ans else None, "rc": r.returncode}
    m = re.search(r"Removed variables\s*:\s*(\d+)", out)
    res["removed_variables"] = int(m.group(1)) if m else None
    if not os.path.exists(prf):
        res["error"] = "no proof file"; res["tail"] = out[-600:]
        return res
    lines = parse(open(prf, "rb").read())
    clauses = [tuple(int(x) for x in lits[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)]
    ck = helpers.RupChecker(clauses)
    added = deleted = not_rup = missing = 0
    empty = False
    first_bad = None
    for kind, c in lines:
        if kind == b"a":
            added += 1
            if not c:
                empty = ck.rup(())
                break
            if not ck.rup(c):
                not_rup += 1
                if first_bad is None:
                    first_bad = [added, list(c)]
            ck.add(c)
        else:
            deleted += 1
            if not ck.delete(c):
                missing += 1
    if not empty and res["answer"] == "UNSATISFIABLE":
        empty = ck.rup(())          # the refutation may end with conflicting units instead of an explicit empty clause
    res.update({"proof_lines": len(lines), "added": added, "deleted": deleted, "not_rup": not_rup, "deleted_missing": missing,
                "refutation_complete": bool(empty), "first_not_rup": first_bad})
    return res


if __name__ == "__main__":
    if len(sys.argv) > 3:
        print(json.dumps(check(sys.argv[1], int(sys.argv[2]), [int(x) for x in sys.argv[3:]])))
    else:
        for fam, seed, args in (("ksat", 71, [50, 260, 3]), ("ksat", 72, [60, 300, 3]), ("ksat", 73, [40, 400, 4]), ("mult", 31, [6]), ("parity", 41, [300])):
            print(json.dumps(check(fam, seed, args)), flush=True)
