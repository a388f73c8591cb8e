#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c15.jsonl
run() { # cfg, env...
  cfg=$1; shift
  env "$@" timeout 150 python tools/kernel_ab.py $cfg 3 --check >> $O/ab_c15.jsonl 2>> $O/ab_c15.err; echo "$cfg $* rc=$?"
}
run cfg2 SIGMA_OT_FILL=70 SIGMA_OT_CPT=3
run cfg2 SIGMA_OT_FILL=80 SIGMA_OT_CPT=3
run cfg2 SIGMA_OT_FILL=80 SIGMA_OT_KEEP=5
run cfg2 SIGMA_OT_FILL=80 SIGMA_OT_KEEP=4
run cfg2 SIGMA_OT_FILL=80 SIGMA_OT_KEEP=3
run cfg2 SIGMA_OT_FILL=70 SIGMA_OT_KEEP=5
run cfg3 SIGMA_OT_FILL=70
run cfg3 SIGMA_OT_FILL=80
run cfg4 SIGMA_OT_FILL=70
run cfg4 SIGMA_OT_FILL=80
run cfg1 SIGMA_OT_FILL=80
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c15.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], [t for t in d['top'] if ('k_ot_' in t[0])])
P
tail -3 $O/ab_c15.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 150 -k "small or edge or golden or full_size or oversized or stage_prep" > $O/pytest_gpu_c15.log 2>&1; tail -4 $O/pytest_gpu_c15.log; grep -n "^E " $O/pytest_gpu_c15.log | head
