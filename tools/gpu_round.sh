#!/bin/bash
# tools/gpu_round.sh <tag> [what...] -- one gpurun call's worth of evidence, every step under its
# own timeout.  Runs on the B200 box; everything lands in gpurun_out/<tag>_*.
#   what: tests bench launches full kernels  (default: all five)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01v2'
tag=${1:-run}; shift
what=${*:-tests bench launches full kernels}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
for w in $what; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
      echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    sanitize)  # memcheck of the small smoke formulas (slow: tiny inputs only)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${tag}_sanitize.log 2>&1
      echo "sanitize rc=$?"; grep -E "ERROR SUMMARY|Invalid|smoke" gpurun_out/${tag}_sanitize.log | head -8 ;;
    bench)
      timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
      echo "bench rc=$?"; cut -c1-600 gpurun_out/${tag}_bench_cfg2.json ;;
    benchref)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
      echo "benchref rc=$?" ;;
    kernels)   # CUDA-event time of every kernel on all four single-GPU configs
      timeout 900 python tools/compare_ref_gpu.py ${KCFGS:-cfg1 cfg4 cfg2 cfg3} --no-ref --out gpurun_out/${tag}_engine_kernels.json > gpurun_out/${tag}_engine_kernels.log 2>&1
      echo "kernels rc=$?"; tail -5 gpurun_out/${tag}_engine_kernels.log ;;
    launches)  # the ncu launch list of the bench workload (cold-cache, serialised: shares only)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
          python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_cfg2.log 2>&1
      echo "launches rc=$?"
      python tools/launch_summary.py gpurun_out/${tag}_launches_cfg2.csv > gpurun_out/${tag}_launches_cfg2.txt 2>&1; head -12 gpurun_out/${tag}_launches_cfg2.txt ;;
    full)      # one ncu --set full capture of the streaming kernels + the irregular ones (few launches each)
      timeout 900 ncu --set full --clock-control none --import-source on \
          -k regex:"${NCU_REGEX:-k_scatter|k_part|k_place|k_hist|k_ere_pairs|k_sort_reg|k_mis_clauses|k_gc_copy|k_awaken|k_count\$}" -c ${NCU_COUNT:-24} -f -o /tmp/${tag}_full \
          python tools/profile_run.py ${NCU_CFG:-cfg2} > gpurun_out/${tag}_full.log 2>&1
      echo "full rc=$?"
      ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
      python tools/ncu_digest.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_full_digest.txt 2>&1; head -40 gpurun_out/${tag}_full_digest.txt
      for kn in ${NCU_SRC:-}; do   # per-instruction stall samples of selected kernels (first captured launch each)
        ncu -i /tmp/${tag}_full.ncu-rep --page source --csv --kernel-name regex:$kn --launch-count 1 > gpurun_out/${tag}_src_${kn}.csv 2>/dev/null
      done
      sz=$(stat -c %s /tmp/${tag}_full.ncu-rep 2>/dev/null || echo 0)
      if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp /tmp/${tag}_full.ncu-rep gpurun_out/; fi ;;
  esac
done
