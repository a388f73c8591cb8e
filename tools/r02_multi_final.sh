#!/bin/bash
# Round 2, final multi-GPU call: bash tools/r02_multi_final.sh N   (under gpurun --gpus N) - the default bench line at N ranks
N=${1:-2}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $T bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_cfg2_${N}gpu_final.json 2> $O/bench_cfg2_${N}gpu_final.err; echo "cfg2 x$N rc=$?"; cut -c1-400 $O/bench_cfg2_${N}gpu_final.json; tail -3 $O/bench_cfg2_${N}gpu_final.err
python - <<P
import json
d=json.loads(open('$O/bench_cfg2_${N}gpu_final.json').read().strip().splitlines()[-1])
e=d['e2e']; print(d['n_gpus'], d['value'], d['ms_per_step'], e['ms_per_step'], e.get('serial'), e.get('pcie_probe'))
P
