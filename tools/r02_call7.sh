#!/bin/bash
# Round 2, GPU call 7: round-2 tests + drop-in, cfg5 (small check, then full size on 1 GPU), ncu launch list + full captures
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 400 python -m pytest tests/test_zy_gpu_round2.py tests/test_zzz_gpu_dropin.py -q -m gpu --timeout 180 > $O/pytest_gpu_c7.log 2>&1; tail -5 $O/pytest_gpu_c7.log; grep -n "^E " $O/pytest_gpu_c7.log | head -10
timeout 200 python bench.py --workload cfg5 --scale 0.02 --steps 2 --warmup 1 > $O/bench_cfg5_x002.json 2> $O/bench_cfg5_x002.err; echo "cfg5 small rc=$?"; cut -c1-500 $O/bench_cfg5_x002.json; tail -3 $O/bench_cfg5_x002.err
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 1 > $O/bench_cfg5_1gpu.json 2> $O/bench_cfg5_1gpu.err; echo "cfg5 full rc=$?"; cut -c1-900 $O/bench_cfg5_1gpu.json; tail -3 $O/bench_cfg5_1gpu.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-ref-gpu --pipeline 0 > $O/launches_cfg2.log 2>&1
echo "launches rc=$?"; python tools/launch_summary.py $O/launches_cfg2.csv > $O/launches_cfg2.txt 2>&1; head -14 $O/launches_cfg2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ot_count|k_ot_part2|k_ot_lithist|k_ot_place|k_ot_col|k_ot_bscan|k_mis_round|k_ere_pairs|k_ere_bloom|k_count\$" -c 40 -f -o /tmp/c7_full_cfg2 \
    python tools/profile_run.py cfg2 > $O/full_cfg2.log 2>&1
echo "full cfg2 rc=$?"
ncu -i /tmp/c7_full_cfg2.ncu-rep --page raw --csv > $O/full_cfg2_raw.csv 2>/dev/null
python tools/ncu_digest.py $O/full_cfg2_raw.csv > $O/full_cfg2_digest.txt 2>&1; head -30 $O/full_cfg2_digest.txt
python tools/ncu_digest.py $O/full_cfg2_raw.csv --json > $O/ncu_traffic_cfg2.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ve_phase1|k_sub|k_ve_phase3|k_mis_round|k_mis_push|k_mis_clauses" -c 24 -f -o /tmp/c7_full_cfg3 \
    python tools/profile_run.py cfg3 > $O/full_cfg3.log 2>&1
echo "full cfg3 rc=$?"
ncu -i /tmp/c7_full_cfg3.ncu-rep --page raw --csv > $O/full_cfg3_raw.csv 2>/dev/null
python tools/ncu_digest.py $O/full_cfg3_raw.csv > $O/full_cfg3_digest.txt 2>&1; head -30 $O/full_cfg3_digest.txt
python tools/ncu_digest.py $O/full_cfg3_raw.csv --json > $O/ncu_traffic_cfg3.json 2>/dev/null
rm -f $O/full_cfg2_raw.csv.tmp
