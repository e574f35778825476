#!/usr/bin/env python
"""Markdown summary of ncu --set full reports (one kernel launch each): the metrics profiles/*.md quote.
usage: ncu_summary.py label=path.ncu-rep ... > profiles/ncu_summary_rNN.md"""
import csv
import subprocess
import sys

M = [("duration", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
     ("regs/thread", "launch__registers_per_thread"), ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("FP64 pipe % (sm__inst_executed_pipe_fp64)", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     ("tensor pipe cycles active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
     ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("L1 data-pipe wavefronts % (l1tex__data_pipe_lsu_wavefronts)", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
     ("active lanes per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
     ("warp instructions", "smsp__inst_executed.sum"),
     ("DFMA thread-instr per cycle (GPU)", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"),
     ("DMUL thread-instr per cycle (GPU)", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"),
     ("DADD thread-instr per cycle (GPU)", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"),
     ("local loads (warp instr)", "smsp__sass_inst_executed_op_local_ld.sum"), ("local stores (warp instr)", "smsp__sass_inst_executed_op_local_st.sum"),
     ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
     ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
     ("DRAM throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("SM active cycles avg", "sm__cycles_active.avg"), ("SM active cycles max", "sm__cycles_active.max")]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return r[0], r[1], r[2:]


cols = []
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    h, u, rows = raw(path)
    for r in rows:
        d = {k: (v, un) for k, v, un in zip(h, r, u)}
        cols.append((label if len(rows) == 1 else f"{label} #{rows.index(r)}", d))
print("| metric | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
print("| kernel | " + " | ".join("`" + c[1]["Kernel Name"][0].split("(")[0][-60:] + "`" for c in cols) + " |")
for name, key in M:
    vals = []
    for _, d in cols:
        if key not in d:
            vals.append("-")
            continue
        v, un = d[key]
        try:
            f = float(v)
            v = f"{f:.4g}" if abs(f) < 1e6 else f"{f:.4e}"
        except ValueError:
            pass
        vals.append(f"{v} {un}".strip())
    print(f"| {name} | " + " | ".join(vals) + " |")
# executed FP64 flops = 2 DFMA + DMUL + DADD (thread instructions per cycle x elapsed cycles)
vals = []
for _, d in cols:
    try:
        pc = 2 * float(d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"][0]) + \
             float(d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"][0]) + \
             float(d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"][0])
        cyc = float(d["sm__cycles_elapsed.max"][0])
        t = float(d["gpu__time_duration.sum"][0])
        un = d["gpu__time_duration.sum"][1]
        t *= {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}.get(un, 1e-9)
        vals.append(f"{pc * cyc:.4e} flop = {pc * cyc / t * 1e-12:.2f} TFLOP/s under ncu ({100 * pc / (148 * 128):.1f} % of 148 x 64 DFMA/clk)")
    except (KeyError, ValueError):
        vals.append("-")
print("| executed FP64 flops (2 DFMA + DMUL + DADD) | " + " | ".join(vals) + " |")
# stall reasons
vals = []
for _, d in cols:
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(v[0])) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k}
    tot = sum(st.values()) or 1
    vals.append(", ".join(f"{k} {100 * x / tot:.0f}%" for k, x in sorted(st.items(), key=lambda t: -t[1])[:5]))
print("| warp stall samples (top 5) | " + " | ".join(vals) + " |")
