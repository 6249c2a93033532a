#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep (needs -lineinfo builds and
`--import-source on` captures): ncu -i REP --page source --csv --print-source cuda,sass, aggregated by (file, line).
usage: tools/ncu_lines.py REP [top N]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = defaultdict(lambda: [0, 0, 0, 0, ""])     # instr, thread instr, samples, long_sb samples, text
fname, hdr, line_no, line_txt = None, None, None, ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":
        line_no, line_txt = r[0], r[1].strip()
    d = dict(zip(hdr[2:], r[2:]))
    def num(k):
        try:
            return int(d.get(k) or 0)
        except ValueError:
            return 0
    a = agg[(fname, int(line_no))]
    a[0] += num("Instructions Executed"); a[1] += num("Thread Instructions Executed"); a[2] += num("# Samples"); a[3] += num("stall_long_sb")
    a[4] = line_txt
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("file:line            instr    %instr  thr/instr  %samples  %long_sb | source")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} {a[0]:10d} {100 * a[0] / tot_i:6.1f}% {a[1] / max(a[0], 1):8.1f} {100 * a[2] / tot_s:8.1f}% {100 * a[3] / tot_s:7.1f}% | {a[4][:100]}")
