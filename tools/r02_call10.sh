#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c10.jsonl
timeout 200 python tools/kernel_ab.py cfg2 5 --check >> $O/ab_c10.jsonl 2>> $O/ab_c10.err
timeout 200 python tools/kernel_ab.py cfg3 3 --check >> $O/ab_c10.jsonl 2>> $O/ab_c10.err
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c10.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], d['top'][:8])
P
for k in 4 6; do
  timeout 300 python bench.py --steps 12 --warmup 3 --pipeline $k --no-secondary --no-ref-gpu --no-cpu-baseline > $O/bench_pipe$k.json 2> $O/bench_pipe$k.err
  python - <<P
import json
d=json.load(open('gpurun_out/r02/bench_pipe$k.json'))
print('pipeline $k', d['ms_per_step'], {x:d['e2e'][x] for x in ('ms_per_step','serial','pipelined','pcie_probe')})
P
done
timeout 1100 python -m pytest tests -q -m gpu --timeout 180 > $O/pytest_gpu_c10.log 2>&1; tail -6 $O/pytest_gpu_c10.log; grep -n "^E " $O/pytest_gpu_c10.log | head -10
