#!/bin/bash
# Turns what one GPU-box visit (tools/gpu_round2.sh <tag> ...) left under gpurun_out/<tag>/ into the tracked summaries under
# profiles/ (prefix <tag>_): bench lines, GPU test log, launch list, ncu --set full summaries of every captured kernel with
# stall-ranked functions and lines, sanitizer logs; regenerates profiles/ncu_traffic.json from the same capture.
# Usage: bash tools/make_profiles.sh <tag> [pairs in the profiled launches, default 250000]
TAG=$1; PAIRS=${2:-250000}
IN=gpurun_out/$TAG; OUT=profiles
for f in bench.json bench_ref.json pytest_gpu.log launches.csv launch_summary.txt tmpfs_write_bench.log smoke.log \
         sanitizer_memcheck_smoke.log sanitizer_memcheck_golden.log sanitizer_racecheck_smoke.log sanitizer_racecheck_golden.log cli_scale.json; do
  [ -s $IN/$f ] && cp $IN/$f $OUT/${TAG}_$f
done
[ -s $IN/bench.log ] && grep -E "^\[bench\]" $IN/bench.log > $OUT/${TAG}_bench.log
[ -s $IN/bench_ref.log ] && grep -E "^\[bench\]" $IN/bench_ref.log > $OUT/${TAG}_bench_ref.log
if [ -s $IN/full_raw.csv ]; then
  python tools/ncu_multi.py $IN/full_raw.csv $PAIRS $OUT/ncu_traffic.json > $OUT/${TAG}_ncu_all_kernels_summary.txt
  rm -rf /tmp/ncu_split_$TAG
  python tools/ncu_split.py $IN/full_src.csv.gz /tmp/ncu_split_$TAG > /dev/null
  for c in /tmp/ncu_split_$TAG/*_src.csv; do
    k=$(basename $c _src.csv)
    python tools/ncu_stalls.py $c > $OUT/${TAG}_${k}_stalls_by_function.txt 2>&1
    python tools/ncu_lines.py $c 40 samples > $OUT/${TAG}_${k}_hot_lines.txt 2>&1
    python tools/ncu_codesize.py $c >> $OUT/${TAG}_${k}_stalls_by_function.txt 2>&1
  done
fi
ls $OUT | grep "^${TAG}_" | wc -l
