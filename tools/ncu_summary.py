#!/usr/bin/env python3
"""Print the key metrics of an .ncu-rep (needs ncu on PATH; no GPU).  usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes_srcunit_tex.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks', 'sm__warps_active.avg.per_cycle_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2:]
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') or ('warp_issue_stalled' in h and h.endswith('.pct'))]
for k in KEYS + sorted(stall):
    if k in hdr:
        i = hdr.index(k)
        print(f'{k:92s} {units[i]:10s}', ' | '.join(v[i] for v in vals))
