"""Static look at a kernel's SASS: total instructions and the loops (backward branches) with their body sizes.
usage: python tools/sass_loops.py <object-or-so> <substring of the mangled kernel name> [--dump START END]"""
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
for name, ins in funcs.items():
    if needle not in name:
        continue
    print(name, "instructions:", len(ins))
    index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?\.?(?:L_x_\d+|0x([0-9a-f]+))", s)
        if m and m.group(1):
            t = int(m.group(1), 16)
            if t <= a and t in index:
                loops.append((index[t], i))
    for s, e in sorted(loops, key=lambda x: x[0] - x[1])[:6]:
        print(f"  loop [{s}..{e}] body {e - s + 1} instructions")
    if "--dump" in sys.argv:
        k = sys.argv.index("--dump")
        for a, s in ins[int(sys.argv[k + 1]) : int(sys.argv[k + 2])]:
            print(f"    {a:05x} {s}")
