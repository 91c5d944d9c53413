"""Counts / prints SASS instructions of one kernel that match a regex.
usage: python tools/sass_grep.py <object> <kernel-name-substring> <regex> [--print]"""
import re
import subprocess
import sys

obj, needle, pattern = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = None
count = {}
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur and needle in cur and pattern.search(m.group(2)):
        count[cur] = count.get(cur, 0) + 1
        if "--print" in sys.argv:
            print(m.group(1), m.group(2).strip())
for k, v in count.items():
    print(v, k[-70:])
