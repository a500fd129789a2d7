"""Summarise `ncu -i X.ncu-rep --page raw --csv` (exported on the GPU box: the .ncu-rep files exceed gpurun_out's size limit) and the matching
`--page source --csv` into the text kept under profiles/.   usage: ncu_csv_summary.py <raw.csv> <source.csv> <cells per launch>"""
import csv, sys, subprocess
raw, src, cells = sys.argv[1], sys.argv[2], float(sys.argv[3])
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:160])
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); vals[k] = r[i]
            print("  %-80s %18s %s" % (k, r[i], units[i]))
    st = sorted(((float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]), reverse=True)
    print("  stall reasons (warps per issue-active cycle): " + "  ".join("%s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for v, h in st[:8]))
    try:
        def gb(k):
            v = float(vals[k].replace(",", "")); u = units[hdr.index(k)]
            return v * {"Gbyte": 1., "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1.)
        t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        print("  dram traffic per launch: %.3f GB = %.1f bytes per interior cell; warp instructions per 32 cells: %.0f" % (
            t, t * 1e9 / cells, float(vals["smsp__inst_executed.sum"].replace(",", "")) / (cells / 32)))
    except Exception as e:
        print("  (traffic summary failed: %s)" % e)
print()
print(subprocess.run([sys.executable, __file__.replace("ncu_csv_summary.py", "ncu_source_mix.py"), src, str(int(cells))], capture_output=True, text=True).stdout)
