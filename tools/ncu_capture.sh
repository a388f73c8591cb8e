#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <count> <cfg> [skip]
# Full ncu capture of selected kernels of one simplification; keeps gpurun_out small: the raw
# metric page always comes back as CSV, the .ncu-rep only when it is below 40 MB.
name=$1; regex=$2; count=$3; cfg=$4; skip=${5:-0}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f -o /tmp/$name \
    python tools/profile_run.py $cfg > gpurun_out/$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/$name.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ]; then cp /tmp/$name.ncu-rep gpurun_out/; fi
echo "$name: rep $sz bytes"
