"""Static view of a kernel's innermost hot loop: dumps the SASS between two addresses and counts instructions per pipe.
usage: python tools/sass_hot.py <object> <kernel-name-substring> [--loop N]   (N = index into the loops sorted by size)
Pipes (B300_MICROARCH.md): IMAD* -> fma; FLO/POPC/MUFU -> xu; LDS/STS/LDG/STG/LDGSTS -> lsu; BRA/BSSY/BSYNC/... -> ctrl;
everything else integer -> alu."""
import re
import subprocess
import sys

obj, needle = sys.argv[1], sys.argv[2]
text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))


def pipe(op):
    op = op.split(".")[0]
    if op in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2"):
        return "fma"
    if op in ("FLO", "POPC", "MUFU", "BREV"):
        return "xu"
    if op in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDL", "STL", "LD", "ST", "ATOMS", "ATOMG", "RED", "LDSM"):
        return "lsu"
    if op in ("BRA", "BSSY", "BSYNC", "EXIT", "NOP", "WARPSYNC", "BAR", "CALL", "RET", "BREAK", "DEPBAR", "LDGDEPBAR", "YIELD"):
        return "ctrl"
    if op.startswith("U") or op in ("S2R", "S2UR", "R2UR", "LDC", "LDCU", "CS2R", "SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"):
        return "other"
    return "alu"


for name, ins in funcs.items():
    if needle not in name:
        continue
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    counts = {}
    for a, s_ in ins:
        if lo <= a <= hi:
            t = s_.split()
            op = t[1] if t[0].startswith("@") else t[0]
            counts[pipe(op)] = counts.get(pipe(op), 0) + 1
            if "--quiet" not in sys.argv:
                print(f"{a:05x} {pipe(op):5s} {s_}")
    print(counts, "total", sum(counts.values()))
