"""Summarise `nvcc -Xptxas -v` output: registers / spills per kernel entry (demangled, filtered by a substring)."""
import re, subprocess, sys
t = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for b in t.split("ptxas info    : Compiling entry function")[1:]:
    name = b.split("'")[1]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat not in dem:
        continue
    short = re.sub(r"hb::\(anonymous namespace\)::|hb::", "", dem.split("(CUtensorMap")[0])
    regs = re.search(r"Used (\d+) registers", b)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    print(short[:150], "| regs", regs.group(1) if regs else "?", "| stack/st/ld", sp.groups() if sp else None)
