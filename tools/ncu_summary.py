"""Prints the key metrics of an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
data = [r for r in rows[2:] if len(r) == len(hdr)]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warps_issue_stalled',
        'smsp__average_warp_latency_per_inst_issued', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe', 'sm__inst_executed_pipe', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'sm__throughput.avg.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'local']
skip = ['per_second', 'pct_of_peak_sustained_elapsed', '.max', '.min']
for i, h in enumerate(hdr):
    if any(k in h for k in keys) and not any(k in h for k in skip if 'dram' not in h or k != 'pct_of_peak_sustained_elapsed'):
        vals = [d[i] for d in data]
        if all(v in ('0', '') for v in vals) and 'stalled' in h:
            continue
        print(f"{h:100s} {vals}")
