#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c29.jsonl
run() { cfg=$1; shift; env "$@" timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c29.jsonl 2>> $O/ab_c29.err; echo "$cfg $* rc=$?"; }
run cfg2 SIGMA_X=0
run cfg4 SIGMA_X=0
run cfg3 SIGMA_X=0
run cfg1 SIGMA_X=0
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c29.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d.get('md5_ordered','')[:8], d['top'][:8])
P
tail -3 $O/ab_c29.err
