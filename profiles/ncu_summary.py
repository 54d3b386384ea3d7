#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text report:
   python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("== raw metrics:", rep)
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print("%-80s %-14s %s" % (h, units[i], vals[i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv",
                      "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
S, IE, AT, SRC = ix["# Samples"], ix["Instructions Executed"], \
    ix["Avg. Threads Executed"], ix["Source"]
tot = ser = tot_i = ser_i = 0
ops = {}
nsass = 0
for r in rows[2:]:
    try:
        s, ie, at = int(r[S]), int(r[IE]), float(r[AT])
    except (ValueError, IndexError):
        continue
    nsass += 1
    tot += s
    tot_i += ie
    if at <= 2.0 and ie > 0:
        ser += s
        ser_i += ie
    parts = r[SRC].split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "")
    op = op.split(".")[0]
    o = ops.setdefault(op, [0, 0])
    o[0] += s
    o[1] += ie
print("\n== source page: %d SASS instructions" % nsass)
print("stall samples %d; in code run by <=2 threads (serial LM algebra): "
      "%.1f%%" % (tot, 100.0 * ser / max(tot, 1)))
print("warp instructions %d; serial %.1f%%" % (tot_i, 100.0 * ser_i / max(tot_i, 1)))
print("%-10s %12s %16s" % ("opcode", "samples", "warp-inst"))
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][1])[:18]:
    print("%-10s %12d %16d" % (k, v[0], v[1]))
