#!/usr/bin/env python
"""Static SASS instruction counts per source line / file of one kernel (nvdisasm --print-line-info output).
usage: tools/sass_lines.py all.sass kernel-name-substring [top N]"""
import collections
import re
import sys

path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
inside = False
cur = ("?", 0)
per_line, per_file = collections.Counter(), collections.Counter()
inline_chain = collections.Counter()
total = 0
for line in open(path, errors="replace"):
    if line.startswith(".text."):
        inside = want in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        per_line[cur] += 1
        per_file[cur[0]] += 1
        total += 1
print(f"{total} instructions")
for f, n in per_file.most_common(12):
    print(f"  {n:7d} {n / total * 100:5.1f}%  {f}")
print()
for (f, l), n in per_line.most_common(top):
    print(f"  {n:6d} {n / total * 100:5.1f}%  {f}:{l}")

# per function: attribute each line to the nearest preceding `__device__` / `__global__` definition in its file
import bisect
import os

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zyg_b200", "csrc", "device")
defs = {}
for f in per_file:
    p = os.path.join(root, f)
    if not os.path.exists(p):
        continue
    starts, names = [], []
    for i, text in enumerate(open(p, errors="replace"), 1):
        m = re.search(r"(?:__device__|__global__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
        if m and not text.lstrip().startswith("//"):
            starts.append(i)
            names.append(m.group(1))
    defs[f] = (starts, names)
per_fn = collections.Counter()
for (f, l), n in per_line.items():
    if f in defs and defs[f][0]:
        k = bisect.bisect_right(defs[f][0], l) - 1
        per_fn[(f, defs[f][1][k] if k >= 0 else "?")] += n
    else:
        per_fn[(f, "?")] += n
print()
for (f, fn), n in per_fn.most_common(top):
    print(f"  {n:6d} {n / total * 100:5.1f}%  {f}: {fn}")
