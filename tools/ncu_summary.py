#!/usr/bin/env python3
"""Summarise an ncu report (read on the CPU box): key raw metrics + per-opcode dynamic
instruction mix + hottest stall sites.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [ticks_per_launch warps]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ticks = float(sys.argv[2]) if len(sys.argv) > 2 else 1000.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_branch_targets_threads_divergent.sum"]
for k in keys:
    if k in m:
        print(f"{k} = {m[k][0]} {m[k][1]}")
for k in sorted(m):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(m[k][0])
        if v >= 0.03:
            print(f"{k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} = {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
iA, iE, iS = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
grid = int(m["launch__grid_size"][0]) * int(m["launch__block_size"][0]) / 32.0
W = grid * ticks
tot, byop, samples = 0, collections.Counter(), []
for r in rows[2:]:
    e = int(r[iE])
    tot += e
    t = r[iA].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0].rstrip(";")
    byop[op] += e
    samples.append((int(r[iS]), r[iA].strip(), e / W))
print(f"warp-instructions per tick per warp = {tot / W:.1f}  (static SASS instructions: {len(rows) - 2})")
print("dynamic mix per tick:", ", ".join(f"{op} {c / W:.1f}" for op, c in byop.most_common(24)))
print("hottest sampled instructions:")
for s, a, e in sorted(samples, reverse=True)[:14]:
    print(f"  {s:7d} samples  exec/tick {e:5.2f}  {a}")
