"""Where a kernel's warps wait: the SASS lines with the most stall samples of an `ncu --set full --import-source on`
capture, with the dominant stall reasons of each line, and the stall totals of the kernel.
usage: python tools/ncu_stalls.py X.ncu-rep <kernel index in the report, 0-based> <pixels (or samples) per launch> [lines=16]
The report is read with `ncu -i X.ncu-rep --page source --csv --print-source sass` (run from a directory that does not hold
the sources, or ncu interleaves them)."""
import csv
import os
import subprocess
import sys
import tempfile

report, kernel, units = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 16
text = subprocess.run(["ncu", "-i", os.path.abspath(report), "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                      text=True, cwd=tempfile.gettempdir()).stdout
rows = list(csv.reader(text.splitlines()))
header = next(r for r in rows if "Instructions Executed" in r)
col = {name: i for i, name in enumerate(header)}
data = [r for r in rows if len(r) == len(header) and r[col["Instructions Executed"]].isdigit()]
# every kernel of the report is listed (ncu prints each twice): a kernel starts where the address falls back
starts = [0] + [i for i in range(1, len(data)) if int(data[i][col["Address"]], 16) < int(data[i - 1][col["Address"]], 16)]
segments = [data[a:b] for a, b in zip(starts, starts[1:] + [len(data)])]
segment = segments[2 * kernel]
steps = units / 32.0  # warp-level steps
samples = sum(int(r[col["# Samples"]]) for r in segment)
executed = sum(int(r[col["Instructions Executed"]]) for r in segment)
print(f"{len(segment)} SASS lines, {executed / steps:.1f} warp instructions per 32 units, {samples} stall samples")
reasons = [h for h in header if h.startswith("stall_") and "Not Issued" not in h]
totals = {h: sum(int(r[col[h]] or 0) for r in segment) for h in reasons}
print("  ".join(f"{h[6:]} {100.0 * v / samples:.1f}%" for h, v in sorted(totals.items(), key=lambda x: -x[1]) if v > 0.01 * samples))
order = sorted(range(len(segment)), key=lambda i: -int(segment[i][col["# Samples"]]))[:top]
for i in sorted(order):
    r = segment[i]
    n = int(r[col["# Samples"]])
    why = {h[6:]: int(r[col[h]] or 0) for h in reasons}
    why = ", ".join(f"{k} {v}" for k, v in sorted(why.items(), key=lambda x: -x[1]) if v > 0.1 * n)
    print(f"{i:6d} {100.0 * n / samples:5.2f}%  x{int(r[col['Instructions Executed']]) / steps:5.2f}  {r[col['Source']][:60]:60s} {why}")
