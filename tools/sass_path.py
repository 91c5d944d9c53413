"""Static estimate of a pixel loop's straight path: the shortest control-flow path around the smallest loop that contains
an instruction matching --via (default FLO: only regular-mode samples compute a Golomb parameter), forced through such an
instruction.  Rare paths (flushes, escapes, run mode) only ever add instructions, so the shortest cycle through the
regular-mode code is what a warp executes when none of its lanes needs one of them.  Prints the path with the pipe of
every instruction (tools/sass_hot.py's classification) and the per-pipe totals.

usage: python tools/sass_path.py <object-or-so> <substring of the mangled kernel name> [--via REGEX] [--quiet]
       [--passes N]   N = how many --via instructions the path must pass (3 for three-component pixels)"""
import heapq
import re
import subprocess
import sys


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


def load(obj):
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
    return funcs


def opcode(s):
    t = s.split()
    return t[1] if t[0].startswith("@") else t[0]


def successors(ins, index):
    """(next instruction indices) of instruction i; BSYNC falls through (the reconvergence point follows it in layout)."""
    out = {}
    for i, (a, s) in enumerate(ins):
        op = opcode(s)
        predicated = s.startswith("@")
        nxt = []
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", s.split("BRA", 1)[1])
            target = index.get(int(m.group(1), 16)) if m else None
            if target is not None:
                nxt.append(target)
            if predicated or "BRA.U" in op and re.search(r"UP\d", s) or re.search(r"BRA(\.\w+)*\s+!?U?P\d", s):
                nxt.append(i + 1)
        elif op.startswith("EXIT") and not predicated:
            nxt = []
        elif op.startswith("RET") or op.startswith("BRX"):
            nxt = []
        else:
            nxt.append(i + 1)
        out[i] = [n for n in nxt if 0 <= n < len(ins)]
    return out


def shortest(succ, sources, target_set, allowed):
    """Dijkstra with unit weights from `sources` (dict node -> cost); returns (cost, path) to the first node of target_set."""
    dist, prev = dict(sources), {}
    heap = [(c, n) for n, c in sources.items()]
    heapq.heapify(heap)
    while heap:
        d, n = heapq.heappop(heap)
        if d > dist.get(n, 1 << 30):
            continue
        if n in target_set and d > 0:
            path = [n]
            while path[-1] in prev:
                path.append(prev[path[-1]])
            return d, path[::-1]
        for m in succ[n]:
            if m not in allowed:
                continue
            if d + 1 < dist.get(m, 1 << 30):
                dist[m] = d + 1
                prev[m] = n
                heapq.heappush(heap, (d + 1, m))
    return None, None


def main():
    obj, needle = sys.argv[1], sys.argv[2]
    via = re.compile(sys.argv[sys.argv.index("--via") + 1]) if "--via" in sys.argv else re.compile(r"^FLO\b(?!\.U32)")
    passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else 1
    quiet = "--quiet" in sys.argv
    for name, ins in load(obj).items():
        if needle not in name:
            continue
        index = {a: i for i, (a, _) in enumerate(ins)}
        succ = successors(ins, index)
        # loops = backward branches; the pixel loop is the smallest one that contains a --via instruction
        loops = []
        for i, (a, s) in enumerate(ins):
            if opcode(s).startswith("BRA"):
                for t in succ[i]:
                    if t <= i and t != i + 1:
                        loops.append((t, i))
        vias = [i for i, (_, s) in enumerate(ins) if via.search(opcode(s) + " " + s)]
        candidates = [(e - b, b, e) for b, e in loops if sum(1 for v in vias if b <= v <= e) >= passes]
        if not candidates:
            print(name, "no loop with", via.pattern)
            continue
        _, head, tail = min(candidates)
        allowed = set(range(head, tail + 1))
        inside = [v for v in vias if head <= v <= tail]
        # the path: head -> via -> ... -> via -> tail (the back edge), each leg a shortest path
        best = None

        def extend(cost, path, remaining):
            nonlocal best
            if remaining == 0:
                c, p = shortest(succ, {path[-1]: 0}, {tail}, allowed)
                if c is not None and (best is None or cost + c < best[0]):
                    best = (cost + c, path + p[1:])
                return
            for v in inside:
                if v in path:
                    continue
                c, p = shortest(succ, {path[-1]: 0}, {v}, allowed)
                if c is not None:
                    extend(cost + c, path + p[1:], remaining - 1)

        extend(0, [head], passes)
        if best is None:
            print(name, "no path")
            continue
        cost, path = best
        counts = {}
        for i in path:
            a, s = ins[i]
            pp = pipe(opcode(s))
            counts[pp] = counts.get(pp, 0) + 1
            if not quiet:
                print(f"{a:05x} {pp:5s} {s}")
        short = re.sub(r".*(k_\w+?_tiled)ILi(\d)ELb(\d)E(\w).*", r"\1<\2,\3,\4>", name)
        print(short, "loop", f"[{head}..{tail}]", "straight path", len(path), dict(sorted(counts.items())))


main()
