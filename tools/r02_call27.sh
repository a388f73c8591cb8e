#!/bin/bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1300 python -m pytest tests -q -m gpu --timeout 200 --durations=40 > $O/pytest_gpu_c27.log 2>&1; tail -60 $O/pytest_gpu_c27.log
