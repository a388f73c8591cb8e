#!/bin/bash
# Round 2, final GPU call: the whole -m gpu suite, the default bench line, cfg5 at full size on one GPU
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1300 python -m pytest tests -q -m gpu --timeout 200 > $O/pytest_gpu_final.log 2>&1; tail -5 $O/pytest_gpu_final.log; grep -n "^E " $O/pytest_gpu_final.log | head
timeout 600 python bench.py > $O/bench_cfg2_final.json 2> $O/bench_cfg2_final.err; echo "bench rc=$?"; cut -c1-700 $O/bench_cfg2_final.json; tail -3 $O/bench_cfg2_final.err
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 > $O/bench_cfg5_final.json 2> $O/bench_cfg5_final.err; echo "cfg5 rc=$?"; cut -c1-900 $O/bench_cfg5_final.json; tail -3 $O/bench_cfg5_final.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
