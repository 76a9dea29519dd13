#!/usr/bin/env python
"""Hot spots of one kernel from an ncu report: top SASS instructions by stall samples, and totals per CUDA source line.
  python tools/ncu_hot.py <report.ncu-rep> <kernel-name> [N]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
def page(src):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", src, "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the page lists every kernel of the report: cut out the section of the one asked for
    start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for n, i in enumerate(start):
        if kern in rows[i][1]:
            end = start[n + 1] if n + 1 < len(start) else len(rows)
            return rows[i + 1], rows[i + 2:end]
    raise SystemExit("kernel not found")
h, rows = page("sass")
si, ci, ii, ti = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
lsb, ssb = h.index("stall_long_sb"), h.index("stall_short_sb")
tot = sum(int(r[ci] or 0) for r in rows if len(r) > ci)
toti = sum(int(r[ii] or 0) for r in rows if len(r) > ii)
print("total samples %d, warp instructions %d" % (tot, toti))
top = sorted((r for r in rows if len(r) > ci), key=lambda r: -int(r[ci] or 0))[:N]
for r in top:
    print("%6.2f%%  inst=%10s thr/inst=%5.1f long_sb=%5s short_sb=%5s  %s" % (100.0 * int(r[ci] or 0) / max(tot, 1), r[ii],
          int(r[ti] or 0) / max(int(r[ii] or 0), 1), r[lsb], r[ssb], r[si][:90]))
if "--lines" in sys.argv:
    # the interleaved page: a source line row ("<no>", "<text>", "", ...) followed by the SASS rows compiled from it
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    acc, cur, fname, hdr = collections.OrderedDict(), None, "?", None
    for r in rows:
        if not r: continue
        if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; sc, ic = [k for k, x in enumerate(r) if x == "# Samples"][0], r.index("Instructions Executed"); continue
        if hdr is None or len(r) < ic + 1: continue
        if r[0] != "": cur = (fname, r[0], r[1].strip()[:100]); acc.setdefault(cur, [0, 0]); continue
        if cur is None: continue
        try:
            acc[cur][0] += int(r[sc] or 0); acc[cur][1] += int(r[ic] or 0)
        except ValueError:
            pass
    ts, ti2 = sum(a[0] for a in acc.values()), sum(a[1] for a in acc.values())
    print("\nper source line: samples %d, warp instructions %d" % (ts, ti2))
    for k, a in sorted(acc.items(), key=lambda x: -x[1][0])[:N]:
        print("%6.2f%% smp %6.2f%% inst  %s:%s  %s" % (100.0 * a[0] / max(ts, 1), 100.0 * a[1] / max(ti2, 1), k[0], k[1], k[2]))
