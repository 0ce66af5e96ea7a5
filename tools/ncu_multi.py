"""Key metrics of every kernel in an `ncu --page raw --csv` dump (one launch per row) and DRAM traffic per read pair.
Usage: python tools/ncu_multi.py raw.csv <pairs in the profiled launches> [traffic.json to update]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h, u = rows[0], rows[1]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio']
ki = h.index('Kernel Name')
traffic = {}
seen = set()
for v in rows[2:]:
    name = v[ki].split('(')[0]
    if name in seen:
        continue
    seen.add(name)
    print(f"== {name}")
    for n in want:
        if n in h:
            i = h.index(n)
            print(f"   {n} [{u[i]}] = {v[i]}")
    try:
        tot = sum(float(v[h.index(k)].replace(',', '')) * UNIT[u[h.index(k)]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        t = float(v[h.index('gpu__time_duration.sum')])
        if pairs:
            print(f"   DRAM bytes per read pair = {tot / pairs:.0f}")
            traffic[name] = {"dram_bytes_per_pair": tot / pairs, "profiled_pairs": pairs, "profiled_launch_ms": t,
                             "source": sys.argv[1].split('/')[-2] + '/' + sys.argv[1].split('/')[-1]}
    except Exception as e:
        print("   traffic:", e)
if len(sys.argv) > 3:
    try:
        old = json.load(open(sys.argv[3]))
    except Exception:
        old = {}
    old.update(traffic)
    json.dump(old, open(sys.argv[3], "w"), indent=1, sort_keys=True)
