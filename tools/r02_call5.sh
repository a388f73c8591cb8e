#!/bin/bash
# Round 2, GPU call 5: OT build v2 (default on) - full GPU suite, fallback run with v1 if it fails, A/B timing
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -q -m gpu -x --timeout 180 > $O/pytest_gpu_c5.log 2>&1; rc=$?; tail -6 $O/pytest_gpu_c5.log
if [ $rc -ne 0 ]; then
  grep -n "^E " $O/pytest_gpu_c5.log | head -20
  SIGMA_OT_V2=0 timeout 900 python -m pytest tests -q -m gpu --timeout 180 > $O/pytest_gpu_c5_v1.log 2>&1; tail -12 $O/pytest_gpu_c5_v1.log
fi
rm -f $O/ab_c5.jsonl
for cfg in cfg2 cfg3 cfg4 cfg1; do
  for v in 0 1; do
    SIGMA_OT_V2=$v timeout 200 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c5.jsonl 2>> $O/ab_c5.err
  done
done
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c5.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], d['top'][:7])
P
tail -5 $O/ab_c5.err
