#!/usr/bin/env python
"""Summarise the source page of an `ncu --set full --import-source on` capture: warp-stall samples by reason, dynamic instruction mix,
and samples per 64-instruction window of the SASS (where the time goes inside the kernel).

    python tools/ncu_source_summary.py <report.ncu-rep>
"""
import csv
import io
import subprocess
import sys
from collections import Counter

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
print("kernel:", rows[hi - 1][1] if hi > 0 and len(rows[hi - 1]) > 1 else "?")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("SASS instructions %d, warp instructions executed %d, stall samples %d" % (len(data), inst, tot))
print("\nstall samples by reason (all warps, halo warps included):")
stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
for h, v in sorted(((h, sum(int(r[ix[h]]) for r in data)) for h in stalls), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %8d %5.1f%%" % (h, v, 100. * v / tot))
mix = Counter()
for r in data:
    ops = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    mix[ops[0].split(".")[0]] += int(r[ix["Instructions Executed"]])
print("\ndynamic instruction mix (warp instructions):")
for op, v in mix.most_common(22):
    print("  %-10s %12d %5.1f%%" % (op, v, 100. * v / inst))
fp64 = sum(v for op, v in mix.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("  FP64 total %d (%.1f%%)" % (fp64, 100. * fp64 / inst))
print("\nsamples per window of 64 SASS instructions (nD = FP64 instructions in the window):")
W = 64
for a in range(0, len(data), W):
    blk = data[a:a + W]
    s = sum(int(r[ix["# Samples"]]) for r in blk)
    if s < tot * .002:
        continue
    g = lambda k: sum(int(r[ix[k]]) for r in blk)
    nd = sum(1 for r in blk if [o for o in r[ix["Source"]].split() if not o.startswith("@")][0][0] == "D")
    print("  %5d-%5d %5.1f%%  nD=%2d  wait %5d  barrier %5d  short_sb %5d  math %5d  selected %5d" % (
        a, a + W, 100. * s / tot, nd, g("stall_wait"), g("stall_barrier"), g("stall_short_sb"), g("stall_math"), g("stall_selected")))
