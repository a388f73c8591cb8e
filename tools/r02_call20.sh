#!/bin/bash
# Round 2, GPU call 20: sigma_load32 test, the default bench line, ncu launch list + full captures (cfg2, cfg3)
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 200 python -m pytest tests/test_zy_gpu_round2.py -q -m gpu --timeout 120 -k "32_bit or compact" > $O/pytest_gpu_c20.log 2>&1; tail -3 $O/pytest_gpu_c20.log; grep -n "^E " $O/pytest_gpu_c20.log | head -5
timeout 600 python bench.py > $O/bench_cfg2_v20.json 2> $O/bench_cfg2_v20.err; echo "bench rc=$?"; cut -c1-1500 $O/bench_cfg2_v20.json; tail -3 $O/bench_cfg2_v20.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg2_v20.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-ref-gpu --pipeline 0 > $O/launches_cfg2_v20.log 2>&1
echo "launches rc=$?"; python tools/launch_summary.py $O/launches_cfg2_v20.csv > $O/launches_cfg2_v20.txt 2>&1; head -16 $O/launches_cfg2_v20.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_ot_count|k_ot_part2|k_ot_lithist|k_ot_place|k_ot_col|k_ot_bscan|k_mis_round|k_ere_pairs|k_ere_bloom|k_count\$|k_ve_phase1" -c 44 -f -o /tmp/c20_full_cfg2 \
    python tools/profile_run.py cfg2 > $O/full_cfg2_v20.log 2>&1
echo "full cfg2 rc=$?"
ncu -i /tmp/c20_full_cfg2.ncu-rep --page raw --csv > $O/full_cfg2_v20_raw.csv 2>/dev/null
python tools/ncu_digest.py $O/full_cfg2_v20_raw.csv > $O/full_cfg2_v20_digest.txt 2>&1; head -50 $O/full_cfg2_v20_digest.txt
python tools/ncu_digest.py $O/full_cfg2_v20_raw.csv --json $O/ncu_traffic_cfg2_v20.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_ve_phase1|k_sub|k_ve_phase3|k_mis_round|k_mis_push|k_mis_clauses|k_ot_part2|k_ot_place" -c 30 -f -o /tmp/c20_full_cfg3 \
    python tools/profile_run.py cfg3 > $O/full_cfg3_v20.log 2>&1
echo "full cfg3 rc=$?"
ncu -i /tmp/c20_full_cfg3.ncu-rep --page raw --csv > $O/full_cfg3_v20_raw.csv 2>/dev/null
python tools/ncu_digest.py $O/full_cfg3_v20_raw.csv > $O/full_cfg3_v20_digest.txt 2>&1; head -34 $O/full_cfg3_v20_digest.txt
python tools/ncu_digest.py $O/full_cfg3_v20_raw.csv --json $O/ncu_traffic_cfg3_v20.json
ls -la $O/*v20* | awk '{print $5, $9}'
