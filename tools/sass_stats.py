"""Static SASS statistics of one kernel: opcode histogram of the whole function and of each loop (backward-branch range).
usage: sass_stats.py <cuobjdump -sass dump> <substring of the mangled name> [lo hi]   (lo, hi: hex address range to histogram)"""
import collections, re, sys
txt = open(sys.argv[1]).read().split("Function : ")
body = [b for b in txt[1:] if sys.argv[2] in b.split("\n")[0]]
if not body:
    sys.exit("no such function")
body = body[0]
ins = []
for ln in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
def op(s):
    t = s.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
def hist(lo, hi, title):
    c = collections.Counter(op(s) for a, s in ins if lo <= a <= hi)
    n = sum(c.values())
    f = sum(v for k, v in c.items() if k in FP64)
    print("%s: [%x, %x] %d instructions, FP64 %d" % (title, lo, hi, n, f))
    print("   " + "  ".join("%s %d" % kv for kv in c.most_common(28)))
print("function: %d instructions" % len(ins))
if len(sys.argv) > 4:
    hist(int(sys.argv[3], 16), int(sys.argv[4], 16), "range")
    sys.exit()
loops = []
for a, s in ins:
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", s)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
for lo, hi in sorted(set(loops)):
    if hi - lo > 16 * 40:
        hist(lo, hi, "loop")
calls = [(a, s) for a, s in ins if op(s) in ("CALL", "RET")]
print("calls:", [(hex(a), s[:60]) for a, s in calls][:20])
