#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c30.jsonl
run() { cfg=$1; shift; env "$@" timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c30.jsonl 2>> $O/ab_c30.err; echo "$cfg $* rc=$?"; }
run cfg3 SIGMA_WL_RECORDS=0
run cfg3 SIGMA_WL_RECORDS=1
run cfg2 SIGMA_WL_RECORDS=1
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c30.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d.get('md5_ordered','')[:8], [t for t in d['top'] if ('local' in t[0] or 'records' in t[0])])
P
tail -3 $O/ab_c30.err
