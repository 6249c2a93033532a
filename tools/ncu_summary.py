#!/usr/bin/env python
"""Append one row per kernel of an .ncu-rep capture to profiles/r<round>_ncu_summary.csv (NCU_SUMMARY=<file> picks it; default the
round-2 file, created with the round-1 header when missing).

usage: python tools/ncu_summary.py gpurun_out/r19_t4p.ncu-rep r19_t4p "what this capture shows"
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

import os
OUT = os.environ.get("NCU_SUMMARY", "profiles/r2_ncu_summary.csv")
if not os.path.exists(OUT):
    open(OUT, "w").write(open("profiles/r1_ncu_summary.csv").readline())


def main():
    rep, name, what = sys.argv[1], sys.argv[2], sys.argv[3]
    cols = next(csv.reader(open(OUT)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    at = {h: i for i, h in enumerate(head)}
    with open(OUT, "a", newline="") as f:
        w = csv.writer(f)
        for r in rows[2:]:
            out = [name, what, r[at["Kernel Name"]].replace("vsgpu::<", "")[:70]]
            for c in cols[3:]:
                if c not in at:
                    out.append("")
                    continue
                v, u = r[at[c]], units[at[c]]
                out.append(f"{v} {u}".strip())
            w.writerow(out)
            print(dict(zip(cols, out)))


if __name__ == "__main__":
    main()
