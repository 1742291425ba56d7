"""Compact summary of an `ncu --set full` capture: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/..."""
import csv
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_src_fp64.sum', 'sm__ops_path_tensor_src_fp64.sum.per_second',
        'sm__ops_path_tensor_src_fp64.sum.peak_sustained_elapsed.per_second',
        'sm__ops_path_tensor_src_tf32_dst_fp32.sum', 'sm__ops_path_tensor_src_tf32_dst_fp32.sum.per_second',
        'sm__ops_path_tensor_src_int8.sum', 'sm__ops_path_tensor_src_int8.sum.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg.per_second',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
out = csv.writer(sys.stdout)
out.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(data))])
for k in KEEP:
    if k in hdr:
        i = hdr.index(k)
        out.writerow([k, units[i]] + [r[i] for r in data])
