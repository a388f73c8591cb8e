#!/bin/bash
# Round 2, multi-GPU call: bash tools/r02_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi topo -m > $O/topo_${N}gpu.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_cfg2_${N}gpu.json 2> $O/bench_cfg2_${N}gpu.err; echo "cfg2 x$N rc=$?"; cut -c1-1500 $O/bench_cfg2_${N}gpu.json; tail -3 $O/bench_cfg2_${N}gpu.err
timeout 900 $T bench.py --gpus $N --workload cfg5 --steps 2 --warmup 1 > $O/bench_cfg5_${N}gpu.json 2> $O/bench_cfg5_${N}gpu.err; echo "cfg5 x$N rc=$?"; cut -c1-1200 $O/bench_cfg5_${N}gpu.json; tail -3 $O/bench_cfg5_${N}gpu.err
