#!/bin/bash
# Round 2, GPU call 6: OT v2 with column-scan run starts + light literal histogram: suite, A/B, bench
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1100 python -m pytest tests -q -m gpu --timeout 180 > $O/pytest_gpu_c6.log 2>&1; tail -8 $O/pytest_gpu_c6.log; grep -n "^E " $O/pytest_gpu_c6.log | head -10
rm -f $O/ab_c6.jsonl
for cfg in cfg2 cfg3 cfg4 cfg1; do
  for v in 0 1; do
    SIGMA_OT_V2=$v timeout 200 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c6.jsonl 2>> $O/ab_c6.err
  done
done
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c6.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], d['top'][:8])
P
tail -3 $O/ab_c6.err
timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench_c6.json 2> $O/bench_c6.err; tail -c 1200 $O/bench_c6.json; tail -3 $O/bench_c6.err
