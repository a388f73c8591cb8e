#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c28.jsonl
run() { cfg=$1; shift; env "$@" timeout 150 python tools/kernel_ab.py $cfg 4 --check >> $O/ab_c28.jsonl 2>> $O/ab_c28.err; echo "$cfg $* rc=$?"; }
run cfg2 SIGMA_OT_CARVE=0
run cfg2 SIGMA_OT_CARVE=1
run cfg2 SIGMA_OT_CARVE=0
run cfg2 SIGMA_OT_CARVE=1
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c28.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],3), d.get('md5_ordered','')[:8], [t for t in d['top'] if 'k_ot_part' in t[0]])
P
tail -3 $O/ab_c28.err
