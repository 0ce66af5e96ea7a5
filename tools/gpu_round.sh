#!/bin/bash
# One GPU-box visit: parity tests, a bench line, the ncu launch list and a full capture of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|smoke|bench|benchq|ncu|ncus|ncup ...]   (default: tests bench ncu)
TAG=${1:-run}; shift
WHAT=${*:-tests bench ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
ncu_full() {  # $1 = kernel regex, $2 = output stem
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 1 -f -o $OUT/$2 \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pairs-per-step 250000 > $OUT/ncu_$2.log 2>&1
  echo "ncu full $2 exit $?"
  ncu -i $OUT/$2.ncu-rep --page raw --csv > $OUT/$2_raw.csv 2>/dev/null
  ncu -i $OUT/$2.ncu-rep --page source --csv --print-source cuda,sass > $OUT/$2_src.csv 2>/dev/null
  python tools/ncu_lines.py $OUT/$2_src.csv 60 > $OUT/$2_hot_lines.txt 2>&1
}
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log ;;
smoke)
  timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log ;;
bench)
  timeout 1500 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.log; echo "bench exit $?"
  cat $OUT/bench.json ;;
benchq)
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/benchq.json 2> $OUT/benchq.log; echo "benchq exit $?"
  cat $OUT/benchq.json ;;
ncu)
  # launch list of the same command (short): per-launch durations, cold-cache and serialised
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'probe_kernel|seed_kernel|pair_kernel|align_kernel|rows_kernel|finish_kernel|rescue_kernel' -c 200 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
  echo "ncu launches exit $?"
  ncu_full rows_kernel rows_full
  ncu_full align_kernel_c align_full
  ncu_full pair_kernel pair_full
  ncu_full probe_kernel probe_full ;;
benchse)
  timeout 900 python bench.py --single-end --pairs-per-step 2000000 --steps 5 --warmup 3 --cpu-sample-pairs 500000 > $OUT/bench_se.json 2> $OUT/bench_se.log; echo "benchse exit $?"
  cat $OUT/bench_se.json ;;
ncul)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'probe_kernel|seed_kernel|pair_kernel|align_kernel|rows_kernel|finish_kernel|rescue_kernel' -c 200 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
  echo "ncu launches exit $?" ;;
ncus) ncu_full rows_kernel rows_full ;;
ncua) ncu_full align_kernel_c align_full ;;
ncur) ncu_full rescue_kernel rescue_full ;;
ncupair) ncu_full pair_kernel pair_full ;;
ncup) ncu_full probe_kernel probe_full ;;
esac
done
ls -la $OUT
