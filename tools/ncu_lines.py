#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo).
usage: tools/ncu_lines.py prof.ncu-rep [top N] [stall]   ("stall": rank the lines by stall samples instead of instructions)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
inst = collections.Counter()
stall = collections.Counter()
text = {}
cur_file = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if "Instructions Executed" in r:
        hdr = r
        ii, wi = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= ii or not r[0].isdigit() or r[2] != "-":
        continue  # only the per-line summary rows (Address == '-')
    key = (cur_file, int(r[0]))
    try:
        inst[key] += int(r[ii])
        stall[key] += int(r[wi])
    except ValueError:
        continue
    text[key] = r[1].strip()
T, S = sum(inst.values()), sum(stall.values())
print(f"total warp instructions {T}, stall samples {S}")
ranked = stall if len(sys.argv) > 3 and "stall" == sys.argv[3] else inst
for key, _ in ranked.most_common(top):
    v = inst[key]
    print(f"{v / T * 100:5.1f}% inst {stall[key] / max(S, 1) * 100:5.1f}% stall  {key[0]}:{key[1]:<5d} {text[key][:110]}")
