#!/bin/bash
# Round 2, GPU call 1: parity holes of VERDICT r01 (items 1a-1c, 2)
set -u
mkdir -p gpurun_out/r02 /tmp/d
O=gpurun_out/r02
g++ -O2 -DCNFGEN_MAIN -o build/cnfgen tools/cnfgen.cpp
build/cnfgen ksat 12 /tmp/d/k3.cnf 800 2400 3
# 1b. root cause of the reference's ERE fault: sanitizer on the default launch, then with --ereminthreads=32
timeout 200 compute-sanitizer --print-limit 6 oracle/_ref/ref_driver /tmp/d/k3.cnf /tmp/d/k3.san.sgd -no-lcvefast -quiet > $O/ref_ere_sanitizer_default.log 2>&1
timeout 200 compute-sanitizer --print-limit 6 oracle/_ref/ref_driver /tmp/d/k3.cnf /tmp/d/k3.san2.sgd -no-lcvefast -quiet --ereminthreads=32 > $O/ref_ere_sanitizer_fix.log 2>&1
grep -c "Invalid" $O/ref_ere_sanitizer_default.log $O/ref_ere_sanitizer_fix.log
tail -3 $O/ref_ere_sanitizer_fix.log
# 2. reference GPU build on growing 5-SAT: where does it start to fail, and what faults
for m in 500000 1000000 2000000 4000000 8000000; do
  n=$((m/21)); build/cnfgen ksat 2 /tmp/d/k5.cnf $n $m 5
  timeout 120 oracle/_ref/ref_driver /tmp/d/k5.cnf /tmp/d/k5.sgd -quiet -no-ere -profilegpu > $O/ref_k5_$m.log 2>&1
  echo "ref k5 m=$m rc=$? $(grep -a 'simplify wall' $O/ref_k5_$m.log | tail -1) $(tail -c 300 $O/ref_k5_$m.log | tr '\n' ' ')"
  rm -f /tmp/d/k5.sgd
done
# 1a. drop-in cases, strict
timeout 200 python -m pytest tests/test_zzz_gpu_dropin.py -q --timeout 60 > $O/dropin.log 2>&1; tail -5 $O/dropin.log
# 1b. ERE goldens again with the launch fix; 1c. proofs one process at a time
timeout 420 python tests/golden/make_golden.py --ere-only > $O/golden_ere.log 2>&1; tail -3 $O/golden_ere.log
GOLDEN_PROOF_JOBS=1 timeout 240 python tests/golden/make_golden_proofs.py > $O/golden_proofs.log 2>&1; tail -3 $O/golden_proofs.log
# never measured in round 1: the 3-deep pipeline (e2e)
timeout 150 python bench.py --steps 9 --warmup 3 --pipeline 3 --no-cpu-baseline > $O/bench_cfg2_pipe3.json 2> $O/bench_cfg2_pipe3.err; tail -c 600 $O/bench_cfg2_pipe3.json
