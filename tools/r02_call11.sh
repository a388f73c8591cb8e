#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c11.jsonl
for cfg in cfg2 cfg4 cfg3; do
  for v in 0 1; do
    SIGMA_OT_TMA=$v timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c11.jsonl 2>> $O/ab_c11.err; echo "$cfg tma=$v rc=$?"
  done
done
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c11.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], [t for t in d['top'] if 'place' in t[0] or 'part' in t[0]])
P
tail -3 $O/ab_c11.err
timeout 500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 120 -k "small or edge or oversized or full_size or golden" > $O/pytest_gpu_c11.log 2>&1; tail -4 $O/pytest_gpu_c11.log
