#!/usr/bin/env python
"""Per source line stall samples of an .ncu-rep (needs -lineinfo):
   python profiles/ncu_lines.py rep.ncu-rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv",
                      "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, cur_line, cur_src = "?", 0, ""
hdr = None
acc = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        S = hdr["# Samples"]; IE = hdr["Instructions Executed"]
        AT = hdr["Avg. Threads Executed"]
        continue
    if hdr is None:
        continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1].strip()
        continue
    try:
        s, ie, at = int(r[S]), int(r[IE]), float(r[AT])
    except (ValueError, IndexError):
        continue
    a = acc.setdefault((cur_file, cur_line), [0, 0, 0, cur_src])
    a[0] += s
    a[1] += ie
    if at <= 2.0:
        a[2] += s
tot = sum(a[0] for a in acc.values())
print("total samples", tot)
byfile = {}
for (f, l), a in acc.items():
    b = byfile.setdefault(f, [0, 0])
    b[0] += a[0]; b[1] += a[2]
for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("%-22s samples %7d (%.1f%%)  serial %7d" % (f, b[0], 100.0 * b[0] / tot, b[1]))
print("top lines:")
for (f, l), a in sorted(acc.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("%6d %5.2f%% ser=%6d inst=%11d %s:%s  %s" % (a[0], 100.0 * a[0] / tot, a[2], a[1], f, l, a[3][:70]))
