#!/bin/bash
# per-line stall samples of the two heaviest cfg3 kernels (round 0)
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_ve_phase1_local|k_sub_local" -c 2 -f -o /tmp/c23_src \
    python tools/profile_run.py cfg3 > $O/src_cfg3_c23.log 2>&1
echo "rc=$?"
ncu -i /tmp/c23_src.ncu-rep --page source --csv > $O/src_cfg3_c23.csv 2>/dev/null
ls -la $O/src_cfg3_c23.csv /tmp/c23_src.ncu-rep
head -c 1500 $O/src_cfg3_c23.csv
