#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py launches <launches.csv>          per-kernel totals / shares of a --metrics gpu__time_duration.sum pass
    python tools/ncu_summary.py full <report.ncu-rep> [cells]     key counters of a --set full capture (per launch)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_"


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0][:100]
        a = agg.setdefault(name, [0, 0., r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# %d launches, %.3f ms total device time (ncu serialised, cold-cache: compare shares)" % (sum(a[0] for a in agg.values()), tot / 1e6))
    print("%-100s %6s %12s %8s %10s  grid block" % ("kernel", "n", "total_ms", "share", "avg_us"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-100s %6d %12.3f %7.2f%% %10.1f  %s %s" % (name, a[0], a[1] / 1e6, 100 * a[1] / tot, a[1] / a[0] / 1e3, a[2], a[3]))


def full(path, cells=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data[:1]:
        print("kernel:", r[ki])
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEYS or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")):
                vals[h] = (r[i], units[i])
        for k in KEYS:
            if k in vals:
                print("  %-75s %18s %s" % (k, vals[k][0], vals[k][1]))
        print("  stall reasons (warps per issue-active cycle):")
        st = sorted(((float(v[0]), k) for k, v in vals.items() if k.startswith(STALLS) and v[0] not in ("", "nan")), reverse=True)
        for v, k in st[:8]:
            print("    %-40s %8.3f" % (k[len(STALLS):-len("_per_issue_active.ratio")], v))
        try:
            rd = float(vals["dram__bytes_read.sum"][0]); wr = float(vals["dram__bytes_write.sum"][0])
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            rd *= mult[vals["dram__bytes_read.sum"][1]]; wr *= mult[vals["dram__bytes_write.sum"][1]]
            print("  dram traffic per launch: %.3f GB (read %.3f + write %.3f)" % ((rd + wr) / 1e9, rd / 1e9, wr / 1e9))
            if cells:
                print("  dram bytes per interior cell per launch: %.1f" % ((rd + wr) / cells))
                ins = float(vals["smsp__inst_executed.sum"][0])
                print("  warp instructions per 32 cells: %.0f" % (ins / (cells / 32)))
        except Exception as e:
            print("  (traffic summary unavailable: %s)" % e)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
