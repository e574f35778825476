#!/usr/bin/env python
"""Print the handful of ncu metrics the profiles/ summaries quote, from a .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys, collections
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max']
def rows(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))
def main(rep):
    r = rows(rep, 'raw'); h, u, v = r[0], r[1], r[2]
    for w in WANT:
        if w in h: print(f'  {w} = {v[h.index(w)][:100]} {u[h.index(w)]}')
    st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(x) for k, x in zip(h, v) if k.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in k}
    tot = sum(st.values())
    print('  stall samples:', ', '.join(f'{k} {100*x/tot:.1f}%' for k, x in sorted(st.items(), key=lambda t: -t[1])[:7]))
    r = rows(rep, 'source', ('--print-source', 'sass')); h = r[1]
    iS, iI, iT = h.index('# Samples'), h.index('Instructions Executed'), h.index('Avg. Threads Executed')
    data = [(int(x[iS]), x[1].strip(), int(x[iI]), x[iT], n) for n, x in enumerate(r[2:]) if len(x) > iS]
    ts = sum(d[0] for d in data); ti = sum(d[2] for d in data)
    c = collections.Counter(); s = collections.Counter(); n = collections.Counter(); th = {}
    for d in data: c[d[2]] += d[2]; s[d[2]] += d[0]; n[d[2]] += 1; th[d[2]] = d[3]
    print('  regions (exec count: #instr, share of warp instr, share of samples, threads):')
    for k, x in sorted(c.items(), key=lambda t: -t[1])[:8]: print(f'    {k}: {n[k]} instr, {100*x/ti:.1f}% instr, {100*s[k]/ts:.1f}% samples, {th[k]} thr')
    if len(sys.argv) > 2:
        print('  top stalled instructions:')
        for d in sorted(data, reverse=True)[:int(sys.argv[2])]: print(f'    #{d[4]} {100*d[0]/ts:.1f}% x{d[2]} thr{d[3]} {d[1][:70]}')
main(sys.argv[1])
