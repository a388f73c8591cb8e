"""Debug aid: engine vs oracle proof chunks on one instance (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
import helpers, sgd
from parafrost_b200 import sigma as S
from test_oracle_proof import CASES
from collections import Counter

for name in sys.argv[1:] or ["mult6"]:
    fam, seed, args = CASES[name]
    V, lits, offs = helpers.gen_cnf(fam, seed, args)
    od, ors, _ = helpers.run_oracle(V, lits, offs, proof=True)
    s = S.Simplifier(0, flags=["-proof"])
    s.load(V, lits, offs)
    s.begin()
    reps = []
    while True:
        rep, done = s.round()
        reps.append(rep)
        if done:
            break
    s.finish()
    chunks, cap = s.proof_chunks()
    print(name, "engine rounds", [(r["kind"], r["elected"], r["eliminated"], r["resolvents"], r["units"]) for r in reps])
    print(" oracle rounds", ors.tolist())
    print(" engine chunk sizes", [len(c) for c in chunks], "oracle", [len(c) for c in od.extra["proof"]], "cap", cap, od.extra["proof_cap"])
    for r in range(max(len(chunks), len(od.extra["proof"]))):
        a = Counter(helpers.drat_canonical(chunks[r])) if r < len(chunks) else Counter()
        b = Counter(helpers.drat_canonical(od.extra["proof"][r])) if r < len(od.extra["proof"]) else Counter()
        only_e, only_o = a - b, b - a
        print(f"  chunk {r}: engine {sum(a.values())} lines, oracle {sum(b.values())}; only engine {sum(only_e.values())}, only oracle {sum(only_o.values())}")
        for k, v in list(only_e.items())[:6]:
            print("     +E", k, v)
        for k, v in list(only_o.items())[:6]:
            print("     +O", k, v)
    s.close()
