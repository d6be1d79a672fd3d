"""Print the metrics of an ncu report that matter for the roofline discussion (ncu -i X --page raw --csv | this)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
h, u = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block',
        'launch__occupancy_limit', 'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_', 'sm__pipe_',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_local', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct', 'lts__t_bytes.sum',
        'sm__cycles_active.avg', 'smsp__inst_executed_pipe_']
for r in rows[2:]:
    print('=' * 100)
    for a, b, c in zip(h, u, r):
        if any(k in a for k in keep) and 'pred_on' not in a:
            print(f"{a:95s} {b:14s} {c}")
