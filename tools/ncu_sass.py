#!/usr/bin/env python
"""The SASS of one kernel from an ncu report in address order with executed warp instructions, threads per instruction and samples.
  python tools/ncu_sass.py <report.ncu-rep> <kernel-name>"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, i in enumerate(start):
    if kern in rows[i][1]:
        end = start[n + 1] if n + 1 < len(start) else len(rows)
        h, rows = rows[i + 1], rows[i + 2:end]
        break
si, ci, ii, ti = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
ai = h.index("Address") if "Address" in h else 0
for r in rows:
    if len(r) <= ii: continue
    n = int(r[ii] or 0)
    print("%s %10d %5.1f %6s  %s" % (r[ai][-5:], n, int(r[ti] or 0) / max(n, 1), r[ci], r[si][:100]))
