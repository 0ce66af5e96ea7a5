#!/bin/bash
# One GPU-box visit: parity tests, a bench line, the ncu launch list and a full capture of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|bench|ncu ...]   (default: all three)
TAG=${1:-run}; shift
WHAT=${*:-tests bench ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
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
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'probe_kernel|search_kernel' -c 40 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
  echo "ncu launches exit $?"
  # one full capture of each hot kernel (skip the warm-up launches)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'search_kernel' -s 3 -c 1 -f -o $OUT/search_full \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pairs-per-step 250000 > $OUT/ncu_full_search.log 2>&1
  echo "ncu full search exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'probe_kernel' -s 3 -c 1 -f -o $OUT/probe_full \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pairs-per-step 250000 > $OUT/ncu_full_probe.log 2>&1
  echo "ncu full probe exit $?" ;;
esac
done
ls -la $OUT
