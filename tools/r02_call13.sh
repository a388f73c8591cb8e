#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python tools/check_dropin_proof.py > $O/dropin_proof_check.jsonl 2> $O/dropin_proof_check.err; cat $O/dropin_proof_check.jsonl | cut -c1-400; tail -3 $O/dropin_proof_check.err
SIGMA_FUZZ_SEEDS=500 SIGMA_FUZZ2_SEEDS=120 SIGMA_FUZZ3_SEEDS=120 SIGMA_FUZZ4_SEEDS=30 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -k "fuzz" > $O/pytest_gpu_fuzz_c13.log 2>&1; tail -4 $O/pytest_gpu_fuzz_c13.log; grep -n "^E " $O/pytest_gpu_fuzz_c13.log | head
timeout 1100 python -m pytest tests -q -m gpu --timeout 180 > $O/pytest_gpu_c13.log 2>&1; tail -5 $O/pytest_gpu_c13.log; grep -n "^E " $O/pytest_gpu_c13.log | head
