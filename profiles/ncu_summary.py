#!/usr/bin/env python
"""Condenses ncu output into the text summaries kept under profiles/.
  python profiles/ncu_summary.py launches <launches.csv>
  python profiles/ncu_summary.py kernel <report.ncu-rep>        (needs `ncu` on PATH to read the report)
  python profiles/ncu_summary.py traffic <report.ncu-rep> <fields per launch> <capture name>   -> traffic.json on stdout
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time (ncu gpu__time_duration.sum, cold-cache serialised: compare SHARES)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-28s launches=%4d  total=%10.1f us  share=%5.1f%%" % (k, a[0], a[1] / 1e3, 100 * a[1] / tot))
    print("total %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("## kernel:", r[h.index("Kernel Name")] if "Kernel Name" in h else "?")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print("%-75s %s %s" % (k, r[i], units[i]))


def traffic(path, nfields, capture):
    """DRAM bytes (read + write) per tile field for every kernel in the report: bench.py's roofline.traffic."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(k)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        res.setdefault(name, {"dram_bytes_per_launch": tot, "dram_bytes_per_field": tot / float(nfields),
                              "duration_ms_under_ncu": float(r[h.index("gpu__time_duration.sum")].replace(",", "")) *
                              {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "nsecond": 1e-6}[units[h.index("gpu__time_duration.sum")]]})
    print(json.dumps({"capture": capture, "fields_per_launch": int(nfields), "kernels": res}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
