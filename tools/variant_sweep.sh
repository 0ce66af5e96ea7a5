#!/bin/bash
# Under gpurun: per-kernel durations (ncu launch list) of every variant library in urmap_b200/variants/.
# Usage: bash tools/variant_sweep.sh <tag> [pairs]
TAG=${1:-sweep}; PAIRS=${2:-500000}
OUT=gpurun_out/$TAG; mkdir -p $OUT
cp urmap_b200/liburmb.so $OUT/liburmb_saved.so
for so in urmap_b200/variants/liburmb_*.so; do
  name=$(basename $so .so); name=${name#liburmb_}
  cp $so urmap_b200/liburmb.so
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'probe_kernel|seed_kernel|pair_kernel|align_kernel|rows_kernel|finish_kernel|rescue_kernel' \
      --csv --log-file $OUT/launches_$name.csv python tools/perf_sweep.py --pairs $PAIRS --flags 0 --reps 3 $SWEEP_EXTRA > $OUT/sweep_$name.log 2>&1
  echo "== $name"; tail -1 $OUT/sweep_$name.log
  python tools/launch_summary.py $OUT/launches_$name.csv
done
cp $OUT/liburmb_saved.so urmap_b200/liburmb.so
