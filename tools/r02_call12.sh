#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c12.jsonl
for cfg in cfg2 cfg1 cfg3 cfg4; do
  timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c12.jsonl 2>> $O/ab_c12.err; echo "$cfg rc=$?"
done
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c12.jsonl'):
    d=json.loads(ln)
    print(d['workload'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], d['top'][:6])
P
tail -3 $O/ab_c12.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zy_gpu_round2.py tests/test_zz_gpu_proof.py -q -m gpu -x --timeout 120 -k "small or edge or golden or ere or option or lcvefast or proof" > $O/pytest_gpu_c12.log 2>&1; tail -4 $O/pytest_gpu_c12.log
