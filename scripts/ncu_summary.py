"""Summarise ncu reports (raw page) into a compact table.  usage: ncu_summary.py tag:file ..."""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum']
cols = {}
for arg in sys.argv[1:]:
    tag, f = arg.split(':')
    txt = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    cols[tag] = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
tags = list(cols)
print('| metric | ' + ' | '.join(tags) + ' |')
print('|---|' + '---|' * len(tags))
for w in WANT:
    if all(w in cols[t] for t in tags):
        def fmt(v):
            try:
                x = float(v); return f'{x:.4g}'
            except ValueError:
                return v
        print(f'| {w} [{cols[tags[0]][w][1]}] | ' + ' | '.join(fmt(cols[t][w][0]) for t in tags) + ' |')
