#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/ab_c14.jsonl
run() { # part3 fuse cfg
  SIGMA_OT_PART3=$1 SIGMA_FUSE_SUBVE=$2 timeout 150 python tools/kernel_ab.py $3 3 --check >> $O/ab_c14.jsonl 2>> $O/ab_c14.err; echo "$3 part3=$1 fuse=$2 rc=$?"
}
run 0 0 cfg2; run 1 0 cfg2
run 0 0 cfg3; run 1 0 cfg3; run 1 1 cfg3
run 1 0 cfg1; run 1 1 cfg1
run 0 0 cfg4; run 1 1 cfg4
python - <<'P'
import json
for ln in open('gpurun_out/r02/ab_c14.jsonl'):
    d=json.loads(ln)
    print(d['workload'], d['env'], round(d['ms_device'],2), d['launches'], d['clauses'], d['eliminated'], d.get('md5_ordered','')[:8], [t for t in d['top'] if ('k_ot_part' in t[0] or 'sub' in t[0] or 've_phase1' in t[0])])
P
tail -3 $O/ab_c14.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zy_gpu_round2.py -q -m gpu -x --timeout 150 -k "small or medium or edge or golden or full_size or option_matrix or learnts or fuzz or resident" > $O/pytest_gpu_c14.log 2>&1; tail -4 $O/pytest_gpu_c14.log; grep -n "^E " $O/pytest_gpu_c14.log | head
