#!/bin/bash
# Round 2, GPU call 3: local-copy SUB / BVE phase 1 - parity first, then A/B on cfg3, cfg4, cfg1
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_proof.py -q -m gpu -x --timeout 180 > $O/pytest_gpu_c3.log 2>&1; tail -4 $O/pytest_gpu_c3.log
for cfg in cfg3 cfg4 cfg1; do
  for v in "0 0" "1 0" "0 1" "1 1"; do
    set -- $v
    SIGMA_VE_LOCAL=$1 SIGMA_SUB_LOCAL=$2 timeout 200 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c3.jsonl 2>> $O/ab_c3.err
  done
done
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c3.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], d['top'][:5])
P
