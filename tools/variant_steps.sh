#!/bin/bash
# Under gpurun: back-to-back step time of every variant library in urmap_b200/variants/ (same box, same data).
# Usage: bash tools/variant_steps.sh <tag> [step_sweep args...]
TAG=${1:-vsteps}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
cp urmap_b200/liburmb.so $OUT/liburmb_saved.so
for rep in 1 2; do
for so in urmap_b200/variants/liburmb_*.so; do
  name=$(basename $so .so); name=${name#liburmb_}
  cp $so urmap_b200/liburmb.so
  python tools/step_sweep.py "$@" > $OUT/steps_${name}_$rep.log 2>&1
  echo "== $name (rep $rep)"; grep "ms/step" $OUT/steps_${name}_$rep.log
done
done
cp $OUT/liburmb_saved.so urmap_b200/liburmb.so
