#!/bin/bash
# tools/final_check.sh <tag>: the bench line of the default workload, its key numbers, and the batch workload (config 5) scaled down.
tag=${1:-final}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_cfg2.json")); r = d["roofline"]
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roofline", r["kernel"], r["frac"], r["traffic"], "cpu", d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --workload cfg5 --scale ${CFG5_SCALE:-0.25} --steps 2 --warmup 1 > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; echo "cfg5 rc=$?"
cut -c1-400 gpurun_out/${tag}_bench_cfg5.json; tail -2 gpurun_out/${tag}_bench_cfg5.err
