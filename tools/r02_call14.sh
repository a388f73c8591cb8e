#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c14.jsonl
for v in 0 1; do
for cfg in cfg2 cfg3 cfg4; do
  SIGMA_OT_PART3=$v timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c14.jsonl 2>> $O/ab_c14.err; echo "$cfg part3=$v rc=$?"
done
done
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c14.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], [t for t in d['top'] if 'k_ot' in t[0]])
P
tail -3 $O/ab_c14.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zy_gpu_round2.py -q -m gpu -x --timeout 120 -k "small or edge or golden or fullsize or cfg" > $O/pytest_gpu_c14.log 2>&1; tail -4 $O/pytest_gpu_c14.log
