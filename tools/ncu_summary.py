#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one column per profiled launch, the metrics that matter here."""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]: " + " | ".join(d[i][:48] for d in data))
print("-- warp stall reasons (warps per issue-active cycle) --")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        vals = [float(d[i].replace(',', '')) if d[i] else 0.0 for d in data]
        if max(vals) > 0.05:
            print(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: " + " | ".join(f"{v:.2f}" for v in vals))
