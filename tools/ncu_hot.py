"""Summarises an `ncu --page source --csv` dump: executed warp instructions per opcode and the hottest SASS lines."""
import collections
import csv
import sys

path, px = sys.argv[1], float(sys.argv[2])  # px = warp-level pixel steps (pixels / 32)
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Instructions Executed" in r)
iE, iS = hdr.index("Instructions Executed"), hdr.index("Source")
data = [r for r in rows if len(r) == len(hdr) and r[iE].isdigit()]
tot = sum(int(r[iE]) for r in data)
print("SASS lines", len(data), "total warp instr", tot, "per warp pixel-step", round(tot / px, 1))
ops = collections.Counter()
for r in data:
    t = r[iS].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += int(r[iE])
for op, c in ops.most_common(22):
    print(f"{op:12s} {c / tot * 100:5.1f}%  {c / px:6.1f}/px")
for lo, hi in ((0.5, 1e9), (0.05, 0.5), (0, 0.05)):
    sel = [int(r[iE]) for r in data if lo * px < int(r[iE]) <= hi * px]
    print(f"lines executed {lo}..{hi}/px: {len(sel)} lines, {sum(sel) / px:.1f} instr/px")
if len(sys.argv) > 3:
    for r in data:
        if int(r[iE]) > float(sys.argv[3]) * px:
            print(f"{int(r[iE]) / px:6.2f}  {r[iS]}")
