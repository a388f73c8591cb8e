#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_<what>.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        a = agg.setdefault(name, [0, 0.0, 0.0, r[gi], r[bi]])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms total (ncu: cold cache, serialised - compare SHARES)")
    print(f"{'kernel':42s} {'n':>5s} {'ms':>10s} {'max ms':>9s} {'share':>6s}  grid / block (first launch)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:42s} {a[0]:5d} {a[1]:10.3f} {a[2]:9.3f} {a[1] / tot:6.3f}  {a[3]} / {a[4]}")


if __name__ == "__main__":
    main(sys.argv[1])
