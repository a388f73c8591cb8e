#!/bin/bash
# Round 2, GPU call 4: full GPU suite (new: -lcvefast, resident continuation, trail ranges, rewritten shim), -lcvefast goldens of the reference
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -q -m gpu --timeout 180 > $O/pytest_gpu_c4.log 2>&1; tail -15 $O/pytest_gpu_c4.log
timeout 420 python tests/golden/make_golden.py --fast > $O/golden_fast.log 2>&1; tail -4 $O/golden_fast.log
