"""Static look at a kernel's SASS: per instruction the control fields (stall count, yield, scoreboard set / wait), so that the
issue cycles a single warp needs for a stretch of code can be added up without a GPU.
usage: cuobjdump -sass -fun <mangled> obj.o > k.sass; python scripts/sass_stalls.py k.sass|- [lo_addr hi_addr [-v]]"""
import re
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # quiet under `| head`
from collections import Counter

ins = []
lines = (sys.stdin if sys.argv[1] == "-" else open(sys.argv[1])).read().splitlines()
i = 0
pat = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        hi = int(m2.group(1), 16)
        stall = (hi >> 41) & 0xF
        yld = (hi >> 45) & 1
        wbar = (hi >> 46) & 7
        rbar = (hi >> 49) & 7
        wait = (hi >> 52) & 0x3F
        ins.append((int(m.group(1), 16), m.group(2).strip(), stall, yld, wbar, rbar, wait))
        i += 2
    else:
        i += 1

lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi_a = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 40
sel = [x for x in ins if lo <= x[0] <= hi_a]
tot = sum(x[2] for x in sel)
print(f"{len(sel)} instructions, sum of stall counts {tot}")
ops = Counter()
stall_by = Counter()
for a, t, s, y, w, r, wt in sel:
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    op = op.split(".")[0]
    ops[op] += 1
    stall_by[op] += s
for op, c in ops.most_common():
    print(f"  {op:12s} {c:5d}  stall-sum {stall_by[op]}")
if "-v" in sys.argv:
    for a, t, s, y, w, r, wt in sel:
        print(f"{a:05x} st={s:2d} y={y} w={w} r={r} wait={wt:06b}  {t}")
