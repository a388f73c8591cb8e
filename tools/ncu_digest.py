#!/usr/bin/env python
"""Digest of an `ncu --page raw --csv` export: one line per captured launch with the numbers the
roofline needs (duration, DRAM bytes read+written, DRAM throughput, L2 hit rate, occupancy).
    python tools/ncu_digest.py gpurun_out/x_full_raw.csv > profiles/rNN_ncu_full_<what>.txt"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "rd"),
    ("dram__bytes_write.sum", "wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * m.get(unit, 1)


def to_ms(v, unit):
    m = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
    return v * m.get(unit, 1)


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, units = rows[hi], rows[hi + 1]
    col = {n: i for i, n in enumerate(h)}
    ki = col["Kernel Name"]
    print(f"# {path}")
    print(f"{'kernel':34s} {'ms':>8s} {'dramRd MB':>10s} {'dramWr MB':>10s} {'GB/s':>8s} {'dram%':>6s} {'l2hit%':>7s} {'occ%':>6s} {'sm%':>6s} {'regs':>5s} {'grid':>8s} {'block':>6s}")
    for r in rows[hi + 2:]:
        if len(r) <= ki:
            continue
        def g(name):
            i = col.get(name)
            if i is None or r[i] in ("", "n/a"):
                return None, ""
            try:
                return float(r[i].replace(",", "")), units[i]
            except ValueError:
                return None, ""
        d, du = g("gpu__time_duration.sum")
        rd, ru = g("dram__bytes_read.sum")
        wr, wu = g("dram__bytes_write.sum")
        ms = to_ms(d, du) if d is not None else float("nan")
        rdb = to_bytes(rd, ru) if rd is not None else float("nan")
        wrb = to_bytes(wr, wu) if wr is not None else float("nan")
        gbs = (rdb + wrb) / (ms * 1e-3) / 1e9 if ms == ms and ms > 0 else float("nan")
        rest = []
        for n, _ in WANT[3:]:
            v, _u = g(n)
            rest.append(v if v is not None else float("nan"))
        name = r[ki].split("(")[0][:34]
        print(f"{name:34s} {ms:8.3f} {rdb / 1e6:10.1f} {wrb / 1e6:10.1f} {gbs:8.0f} {rest[0]:6.1f} {rest[1]:7.1f} {rest[2]:6.1f} {rest[3]:6.1f} {rest[4]:5.0f} {rest[5]:8.0f} {rest[6]:6.0f}")


def traffic_json(path, out):
    """--json: mean DRAM bytes per launch and duration per kernel -> profiles/*_ncu_traffic_<cfg>.json (bench.py `traffic`)."""
    import json
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, units = rows[hi], rows[hi + 1]
    col = {n: i for i, n in enumerate(h)}
    agg = {}
    for r in rows[hi + 2:]:
        if len(r) <= col["Kernel Name"]:
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
        try:
            rd = to_bytes(float(r[col["dram__bytes_read.sum"]].replace(",", "")), units[col["dram__bytes_read.sum"]])
            wr = to_bytes(float(r[col["dram__bytes_write.sum"]].replace(",", "")), units[col["dram__bytes_write.sum"]])
        except (ValueError, KeyError):
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += rd + wr
    json.dump({"source": path, "how": "ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, mean per captured launch",
               "kernels": {k: v[1] / v[0] for k, v in agg.items()}}, open(out, "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--json":
        traffic_json(sys.argv[1], sys.argv[3])
    else:
        main(sys.argv[1])
