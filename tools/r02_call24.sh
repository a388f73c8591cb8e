#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c24.jsonl
run() { cfg=$1; shift; env "$@" timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c24.jsonl 2>> $O/ab_c24.err; echo "$cfg $* rc=$?"; }
run cfg3 SIGMA_SUB_THREAD=0
run cfg3 SIGMA_SUB_THREAD=1
run cfg1 SIGMA_SUB_THREAD=1
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c24.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d.get('md5_ordered','')[:8], [t for t in d['top'] if 'sub' in t[0]])
P
tail -3 $O/ab_c24.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 150 -k "small or medium or edge or option_matrix or learnts or fuzz" > $O/pytest_gpu_c24.log 2>&1; tail -4 $O/pytest_gpu_c24.log; grep -n "^E " $O/pytest_gpu_c24.log | head
