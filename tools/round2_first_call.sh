#!/bin/bash
# First GPU call of the next round (one B200, ~6 minutes):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/round2_first_call.sh'
# 1. proof files of the unmodified reference binary, one process at a time (tests/golden/make_golden_proofs.py)
#    -> copy gpurun_out/golden_proof/*.drat.gz + summary.json entries with proof_bytes > 0 into tests/golden/proof/
# 2. the drop-in CLI tests, incl. the cases still marked xfail (answers + -modelverify against the reference binary)
# 3. the full GPU suite, the bench line and the ncu launch list of the same command
set -u
mkdir -p gpurun_out
GOLDEN_PROOF_JOBS=1 timeout 170 python tests/golden/make_golden_proofs.py > gpurun_out/r02_golden_proofs.log 2>&1
timeout 120 python -m pytest tests/test_zzz_gpu_dropin.py -q --runxfail --timeout 60 > gpurun_out/r02_dropin.log 2>&1
timeout 150 python -m pytest tests -q -m gpu -x --timeout 120 > gpurun_out/r02_pytest_gpu.log 2>&1
timeout 90 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err
timeout 120 python bench.py --steps 9 --warmup 3 --pipeline 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg2_pipe3.json 2> gpurun_out/r02_bench_cfg2_pipe3.err
tail -3 gpurun_out/r02_golden_proofs.log gpurun_out/r02_dropin.log gpurun_out/r02_pytest_gpu.log gpurun_out/r02_bench_cfg2.json
