#!/usr/bin/env python
"""Per-kernel table of an .ncu-rep with several kernels (raw page):
   python profiles/ncu_multi.py gpurun_out/prof_kernels.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"),
        ("gpu__time_duration.sum", "time"),
        ("launch__grid_size", "grid"),
        ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("smsp__inst_executed.sum", "inst"),
        ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sass__inst_executed_local_loads", "lld"),
        ("sass__inst_executed_local_stores", "lst")]
stall = [h for h in hdr if "issue_stalled" in h
         and h.endswith("per_issue_active.ratio")]
print("== %s" % rep)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("-" * 78)
    for h, short in cols:
        if h in ix:
            print("%-8s %-14s %s" % (short, units[ix[h]], r[ix[h]][:150]))
    st = sorted(((float(r[ix[h]] or 0), h.split("issue_stalled_")[1]
                  .replace("_per_issue_active.ratio", "")) for h in stall),
                reverse=True)[:6]
    print("stalls/issue: " + ", ".join("%s %.2f" % (b, a) for a, b in st))
