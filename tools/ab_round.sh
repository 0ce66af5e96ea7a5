#!/bin/bash
# Under gpurun: A/B of the variant libraries in urmap_b200/variants/ on the human-scale workload (same box, same data):
# back-to-back step time, per-kernel-class times and a result signature per variant, for URMB_FLAGS 0 and 64.
# Usage: bash tools/ab_round.sh <tag> [variant names...]
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
cp urmap_b200/liburmb.so $OUT/liburmb_saved.so
for name in "$@"; do
  cp urmap_b200/variants/liburmb_$name.so urmap_b200/liburmb.so
  timeout 400 python tools/step_sweep.py --var URMB_FLAGS --values ${VALUES:-0,64} --steps 9 > $OUT/steps_$name.log 2>&1
  echo "== $name"; grep "ms/step" $OUT/steps_$name.log
done
cp $OUT/liburmb_saved.so urmap_b200/liburmb.so
rm -f $OUT/liburmb_saved.so
