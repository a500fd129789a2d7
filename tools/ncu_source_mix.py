"""Dynamic instruction mix and stall summary from `ncu --page source --csv` of one kernel launch.
usage: ncu_source_mix.py <source.csv> [cells]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
out = []
for n, h in enumerate(hi[:1]):
    hdr = rows[h]
    data = rows[h + 1: hi[n + 1] - 1 if n + 1 < len(hi) else None]
    ix = {k: i for i, k in enumerate(hdr)}
    ex = collections.Counter(); smp = collections.Counter(); tot = 0
    for r in data:
        if len(r) < len(hdr): continue
        s = r[ix["Source"]].split()
        if not s: continue
        if s[0].startswith("@"): s = s[1:]
        op = s[0].split(".")[0]
        e = int(r[ix["Instructions Executed"]]); ex[op] += e; tot += e
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
    print("SASS lines %d, warp instructions executed %d%s" % (len(data), tot, " = %.1f per 32 cells" % (tot / (cells / 32)) if cells else ""))
    f64 = sum(v for k, v in ex.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print("FP64 %d (%.1f%%)%s" % (f64, 100. * f64 / tot, " = %.1f per 32 cells" % (f64 / (cells / 32)) if cells else ""))
    for k, v in ex.most_common(40):
        print("  %-10s %12d %5.1f%% %s" % (k, v, 100. * v / tot, "%.1f" % (v / (cells / 32)) if cells else ""))
    st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    S = {k: sum(int(r[ix[k]]) for r in data if len(r) >= len(hdr)) for k in st}
    T = sum(S.values())
    print("stall samples:", "  ".join("%s %.1f%%" % (k[6:], 100. * v / T) for k, v in sorted(S.items(), key=lambda kv: -kv[1])[:10]))
