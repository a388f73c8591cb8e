#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c19.jsonl
run() { cfg=$1; shift; env "$@" timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c19.jsonl 2>> $O/ab_c19.err; echo "$cfg $* rc=$?"; }
run cfg2 SIGMA_X=0
run cfg4 SIGMA_X=0
run cfg3 SIGMA_X=0
run cfg1 SIGMA_X=0
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c19.jsonl'):
    d=json.loads(ln)
    print(d['workload'], round(d['ms_device'],2), d['launches'], d.get('md5_ordered','')[:8], d['top'][:9])
P
tail -3 $O/ab_c19.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zy_gpu_round2.py -q -m gpu -x --timeout 150 -k "small or medium or edge or golden or option_matrix or lcvefast or fuzz" > $O/pytest_gpu_c19.log 2>&1; tail -4 $O/pytest_gpu_c19.log; grep -n "^E " $O/pytest_gpu_c19.log | head
