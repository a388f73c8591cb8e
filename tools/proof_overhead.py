"""Device time of simplify() with and without the DRAT proof stream (flag -proof) on medium instances.
Usage (GPU box): python tools/proof_overhead.py > gpurun_out/proof_overhead.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers
from parafrost_b200 import sigma as S

CASES = {"cfg1_k3_100k": ("ksat", 1, [100000, 426000, 3]), "miter_50k": ("miter", 3, [2000, 50000, 900, 100, 32]), "mult48": ("mult", 4, [48])}
out = {}
for name, (fam, seed, args) in CASES.items():
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    row = {"V": int(V), "C": int(len(offs) - 1), "L": int(len(lits))}
    for tag, flags in (("off", []), ("on", ["-proof"])):
        s = S.Simplifier(0, flags=flags)
        s.load(V, lits, offs)
        ms = []
        for _ in range(4):
            rep = s.simplify()
            ms.append(rep["ms_device"])
        row[tag + "_ms_device"] = round(min(ms[1:]), 3)
        row[tag + "_launches"] = int(rep["kernel_launches"])
        if flags:
            chunks, cap = s.proof_chunks()
            row["proof_bytes"] = sum(len(c) for c in chunks)
            row["proof_lines"] = sum(len(helpers.drat_parse(c)) for c in chunks) if row["proof_bytes"] < 5_000_000 else None
            row["proof_cap"] = cap
        s.close()
    out[name] = row
print(json.dumps(out, indent=1))
