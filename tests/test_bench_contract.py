"""CPU-side checks of bench.py's contract: the reference arm prints exactly one JSON line with the agreed
keys (it times the reference's CPU simplifier, oracle/_ref/parafrost_cpu), and the roofline helper maps
the timer's kernel names to their byte formulas."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_prints_one_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "parafrost_cpu")):
        pytest.skip("reference CPU build not present (make -f oracle/Makefile ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "simplify_literals_per_s" and d["unit"] == "literals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["config"]["workload"] == "cfg1"


def test_roofline_is_the_dominant_kernel_with_engine_counted_bytes():
    b = load_bench()
    assert b.kbase("(k_ot_part<3, 5>)") == "k_ot_part" and b.kbase("void k_sub<4>") == "k_sub" and b.kbase("k_count") == "k_count"
    # {kernel: (ms, launches, algorithmic bytes)} as sigma_kernel_stats returns it
    ks = {"k_mis_round<32>": (9.0, 100, 9.0e9), "(k_ot_part<3, 5>)": (5.0, 5, 5 * 1.596e9), "k_awaken": (4.0, 5, 0.0)}
    r = b.roofline_block(ks, {"hbm_gbs": 6551.4}, "cfg2")
    assert r["kernel"] == "k_mis_round<32>" and r["launches"] == 100 and abs(r["ms_per_launch"] - 0.09) < 1e-9
    assert abs(r["achieved"] - 1000.0) < 0.1 and abs(r["frac"] - 1000.0 / 6551.4) < 1e-3
    assert r["algorithmic_bytes_per_launch"] == 9.0e7
    rows = {x["kernel"]: x for x in r["top_kernels"]}
    assert abs(rows["(k_ot_part<3, 5>)"]["gbs"] - 1596.0) < 0.1 and rows["k_awaken"]["gbs"] is None
    r2 = b.roofline_block({"k_awaken": (4.0, 5, 4e9)}, {}, "cfg2")
    assert r2["peak"] == 6650.0 and "fallback" in r2["peak_source"]


def test_batch_specs_cover_the_config5_range():
    sys.path.insert(0, ROOT)
    from parafrost_b200 import replicas
    specs = replicas.batch_specs(64)
    w = [replicas.spec_weight(s) for s in specs]
    assert len(specs) == 64 and 0.9e6 <= min(w) <= 1.1e6 and 4.0e7 <= max(w) <= 6.0e7
    assert {s[0] for s in specs} == {"ksat", "miter", "multpar"}


def test_traffic_files_are_ordered_by_round_then_capture(tmp_path, monkeypatch):
    """profiles/rNN_ncu_traffic_<cfg>_vMM.json: a later round overrides an earlier one whatever its capture number."""
    b = load_bench()
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "r01_ncu_traffic_cfg2_v41.json").write_text(json.dumps({"kernels": {"k_a": 1.0, "k_b": 10.0}}))
    (prof / "r02_ncu_traffic_cfg2_v7.json").write_text(json.dumps({"kernels": {"k_a": 2.0}}))
    (prof / "r02_ncu_traffic_cfg2_v20.json").write_text(json.dumps({"kernels": {"k_a": 3.0}}))
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    assert b.ncu_traffic("cfg2") == {"k_a": 3.0, "k_b": 10.0}


def test_one_clock_sampler_watches_every_gpu_of_the_job():
    b = load_bench()
    assert b.ClockSampler(range(4)).index == "0,1,2,3" and b.ClockSampler(0).index == "0"
    s = b.ClockSampler(range(2))
    s.p, s.t = type("P", (), {"terminate": lambda self: None})(), type("T", (), {"join": lambda self, timeout=None: None})()
    s.lines = ["0, 1965, 1965, 400.0, Not Active, Not Active, Not Active, Not Active\n", "1, 1950, 1965, 900.0, Not Active, Not Active, Not Active, Active\n"]
    c = s.stop()
    assert c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["samples"] == 2


def test_offsets_go_over_pcie_as_32_bit_words_when_they_fit():
    import numpy as np
    b = load_bench()
    alloc = lambda n, dt: np.empty(int(n), dt)
    o = b.offs32_of(np.array([0, 3, 8], np.uint64), alloc)
    assert o.dtype == np.uint32 and o.tolist() == [0, 3, 8]
    big = np.array([0, 1 << 32], np.uint64)
    assert b.offs32_of(big, alloc) is big
