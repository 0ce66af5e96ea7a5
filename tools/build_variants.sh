#!/bin/bash
# Builds launch-bounds variants of liburmb.so into urmap_b200/variants/ (tuning only; they travel with gpurun).
# Usage: tools/build_variants.sh name:PAIR:ROWS:ALIGN ...
cd "$(dirname "$0")/../urmap_b200/csrc" || exit 1
mkdir -p ../variants
for v in "$@"; do
  IFS=: read -r name p r a <<< "$v"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -w \
       -DURMB_LB_PAIR=$p -DURMB_LB_ROWS=$r -DURMB_LB_ALIGN=$a -shared -o ../variants/liburmb_$name.so \
       $EXTRA_DEFS urmb_kernels.cu urmb_big.cu urmb_api.cu urmb_build.cu urmb_peaks.cu &
done
wait
ls -la ../variants
