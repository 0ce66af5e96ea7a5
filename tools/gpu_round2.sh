#!/bin/bash
# One GPU-box visit of round 2.  Usage (under gpurun): bash tools/gpu_round2.sh <tag> [step ...]
#   tests      pytest -m gpu
#   smoke      __graft_entry__.smoke()
#   tiny       bench.py on a 20 Mb genome with every leg (checks the plumbing in a minute)
#   bench      the default bench line (all configurations, reference binary, CLI)
#   benchq     kernels only (no CPU legs, no other configurations)
#   benchref   the reference arm
#   launches   ncu launch list of a short bench run (every kernel)
#   ncufull    one `ncu --set full` capture of every kernel class of one step (CSV exports made on the box)
#   sanitize   compute-sanitizer memcheck + racecheck on the golden sets
#   ab         per-kernel step times of the current build (tools/step_sweep.py)
#   tmpfs      threads filling a fresh file in /dev/shm (the ceiling of the SAM sink on this box)
#   cliscale   BASELINE's full size through the drop-in next to the reference binary (tools/cli_scale.py, 10 M pairs)
TAG=${1:-run}; shift
WHAT=${*:-tests bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
KREGEX='probe_pair_kernel|probe_kernel|seed_kernel|pair_kernel|align_kernel|rows_kernel|rows_long_kernel|finish_kernel|rescue'
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log ;;
smoke)
  timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log ;;
tiny)
  timeout 900 python bench.py --genome-len 20000000 --pairs-per-step 20000 --config-units-scale 0.02 --steps 3 --cpu-sample-pairs 20000 \
      > $OUT/bench_tiny.json 2> $OUT/bench_tiny.log; echo "tiny exit $?"
  tail -c 1500 $OUT/bench_tiny.log; head -c 600 $OUT/bench_tiny.json; echo ;;
bench)
  URMB_BENCH_UNMATCHED=$OUT/unmatched timeout 1700 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.log; echo "bench exit $?"
  tail -c 3000 $OUT/bench.log; head -c 1200 $OUT/bench.json; echo ;;
benchq)
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > $OUT/benchq.json 2> $OUT/benchq.log; echo "benchq exit $?"
  tail -c 800 $OUT/benchq.log ;;
benchcfg)
  URMB_DEBUG=${BENCH_DEBUG:-} timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BENCHCFG_ARGS > $OUT/benchcfg.json 2> $OUT/benchcfg.log; echo "benchcfg exit $?"
  grep -E "config|main workload" $OUT/benchcfg.log; grep "rescue rounds" $OUT/benchcfg.log | sort | uniq -c | sort -rn | head -12 ;;
benchref)
  timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.log; echo "benchref exit $?"
  tail -c 1000 $OUT/bench_ref.log; cat $OUT/bench_ref.json ;;
launches)
  # launch list of the same command (short): per-launch durations, cold-cache and serialised
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREGEX" -c 400 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > $OUT/ncu_launch_bench.log 2>&1
  echo "ncu launches exit $?"
  python tools/launch_summary.py $OUT/launches.csv > $OUT/launch_summary.txt 2>&1; cat $OUT/launch_summary.txt ;;
ncufull)
  # 250 k pairs per step = one chunk: every step launches each kernel class once, in a fixed order; skip the warm-up
  # steps and capture one launch of every class (finish_kernel is excluded: 0.3 ms)
  # per step and in launch order: probe_pair, pair, align_a, rows, rows_long, align_c, rescue_last (the in-place mate rescue):
  # 7 matching launches; three warm-up steps and the first timed step are skipped, then one launch of every class is captured
  # (URMB_RESCUE_ROUNDS=r adds r x (rescue_scan, rescue_dp) per step: set NCU_SKIP / NCU_COUNT accordingly)
  NK=${NCU_KERNELS:-'probe_pair_kernel|probe_kernel|pair_kernel|align_kernel_a|align_kernel_c|rows_kernel|rows_long_kernel|rescue_last_kernel|rescue_scan_kernel|rescue_dp_kernel'}
  NSKIP=${NCU_SKIP:-28}
  NCOUNT=${NCU_COUNT:-7}
  timeout 1700 ncu --set full --clock-control none --import-source on -k regex:"$NK" -s $NSKIP -c $NCOUNT -f -o $OUT/${NCU_OUT:-full} \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --pairs-per-step ${NCU_PAIRS:-250000} $NCU_BENCH_ARGS > $OUT/ncu_full.log 2>&1
  echo "ncu full exit $?"
  F=${NCU_OUT:-full}
  ncu -i $OUT/$F.ncu-rep --page raw --csv > $OUT/${F}_raw.csv 2>/dev/null
  ncu -i $OUT/$F.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $OUT/${F}_src.csv.gz
  ls -la $OUT/$F.ncu-rep
  rm -f $OUT/$F.ncu-rep ;;   # the CSV exports are what is read afterwards; the report itself would eat the 64 MiB return budget
sanitize)
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > $OUT/sanitizer_memcheck_smoke.log 2>&1
  echo "memcheck smoke exit $?"; tail -4 $OUT/sanitizer_memcheck_smoke.log
  timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or big_capacity or dense_index or first_look or maxix" \
      > $OUT/sanitizer_memcheck_golden.log 2>&1
  echo "memcheck golden exit $?"; tail -4 $OUT/sanitizer_memcheck_golden.log
  URMB_FLAGS=256 timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > $OUT/sanitizer_racecheck_smoke.log 2>&1
  echo "racecheck smoke exit $?"; tail -4 $OUT/sanitizer_racecheck_smoke.log
  timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "golden" \
      > $OUT/sanitizer_racecheck_golden.log 2>&1
  echo "racecheck golden exit $?"; tail -4 $OUT/sanitizer_racecheck_golden.log ;;
cliscale)
  timeout 1700 python tools/cli_scale.py --pairs ${CLI_PAIRS:-10000000} --tabbedout > $OUT/cli_scale.json 2> $OUT/cli_scale.log; echo "cli_scale exit $?"
  tail -3 $OUT/cli_scale.log; head -c 1500 $OUT/cli_scale.json; echo ;;
tmpfs)
  # what the SAM sink can reach on this box: threads filling a fresh 4 GiB file in /dev/shm (tools/tmpfs_write_bench.cpp)
  g++ -O2 -pthread -o /tmp/tmpfs_write_bench tools/tmpfs_write_bench.cpp
  for m in 0 2 1 3; do for t in 1 4 8 16; do /tmp/tmpfs_write_bench /dev/shm/urmb_wbtest $m $t 4; done; done > $OUT/tmpfs_write_bench.log 2>&1
  cat $OUT/tmpfs_write_bench.log ;;
ab)
  timeout 900 python tools/step_sweep.py > $OUT/step_sweep.log 2>&1; echo "step_sweep exit $?"; tail -12 $OUT/step_sweep.log ;;
esac
done
ls -la $OUT
