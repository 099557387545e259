#!/usr/bin/env python3
"""tools/ncu_summary.py <report.ncu-rep> -- key per-kernel metrics of an ncu --set full report (run here, no GPU needed)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'sm__inst_executed.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        # the L1TEX data pipe (the unit the row kernels saturate): wavefronts by source, register write-back, and what reaches L2
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__lsu_writeback_active.sum.pct_of_peak_sustained_elapsed', 'l1tex__lsu_writeback_active_mem_lgds.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum',
        'l1tex__m_xbar2l1tex_read_sectors.sum.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex.sum',
        'lts__t_sectors_srcunit_tex.sum.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_lookup_hit.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'lts__t_sectors.sum.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
for d in data:
    print('----', d[idx['Kernel Name']][:70], 'grid', d[idx['Grid Size']], 'block', d[idx['Block Size']])
    for w in WANT:
        if w in idx:
            print(f"   {w:82s} {d[idx[w]]:>18s} {units[idx[w]]}")
