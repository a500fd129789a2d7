"""Registers / spills per fv_march instantiation: nvcc -Xptxas -v output on stdin -> one line per kernel."""
import re
import sys
cur = None
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        n = m.group(1)
        k = re.search(r"fv_marchINS_\d+(\w+?)I([df])Lb(\d)EEELi(\d+)ELi(\d+)ENS_8MarchCfgILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)EEE", n)
        cur = ("%s<%s,%s> dim=%s lim=%s cfg<%s,%s,%s,%s,%s>" % ((k.group(1), k.group(2), k.group(3), k.group(4), k.group(5)) + k.groups()[5:])) if k else None
        stack = None
    elif cur and "stack frame" in line:
        stack = re.findall(r"(\d+) bytes", line)
    elif cur and "Used" in line:
        r = re.search(r"Used (\d+) registers", line).group(1)
        print("%-60s regs=%s stack=%s spill_st=%s spill_ld=%s" % (cur, r, stack[0], stack[1], stack[2]))
        cur = None
