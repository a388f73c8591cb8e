#!/bin/bash
# Round 2, GPU call 2: why the reference GPU build dies above ~2M clauses; regression suite after the API changes; new bench line
set -u
mkdir -p gpurun_out/r02 /tmp/d
O=gpurun_out/r02
g++ -O2 -DCNFGEN_MAIN -o build/cnfgen tools/cnfgen.cpp
build/cnfgen ksat 2 /tmp/d/k5.cnf 190476 4000000 5
CUDA_LAUNCH_BLOCKING=1 timeout 120 oracle/_ref/ref_driver /tmp/d/k5.cnf /tmp/d/k5.sgd -no-ere --verbose=3 > $O/ref_k5_4M_blocking.log 2>&1
echo "blocking rc=$?"; sed 's/\x1b\[[0-9;]*m//g' $O/ref_k5_4M_blocking.log | tail -12
timeout 420 compute-sanitizer --print-limit 4 oracle/_ref/ref_driver /tmp/d/k5.cnf /tmp/d/k5.sgd -quiet -no-ere > $O/ref_k5_4M_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -a -m 12 "Invalid\|at \|by thread\|Address\|ERROR SUMMARY\|Host Frame: ParaFROST" $O/ref_k5_4M_sanitizer.log | head -24
rm -f /tmp/d/k5.sgd /tmp/d/k5.cnf
timeout 600 python -m pytest tests -q -m gpu -x --timeout 180 > $O/pytest_gpu_c2.log 2>&1; tail -4 $O/pytest_gpu_c2.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 1500 $O/bench_c2.json; tail -5 $O/bench_c2.err
