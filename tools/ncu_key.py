"""Key metrics from an `ncu --page raw --csv` dump. Usage: python tools/ncu_key.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__sass_inst_executed_op_local', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__average_warps_issue_stalled', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for i, n in enumerate(h):
    if any(n.startswith(w) for w in want):
        if 'stalled' in n:
            try:
                if float(v[i]) < 0.2: continue
            except ValueError: pass
        print(f"{n} [{u[i]}] = {v[i]}")
