#!/usr/bin/env python
"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file ...`).
usage: tools/launch_shares.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void zygpu::<unnamed>::", "").replace("zygpu::<unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"{sys.argv[1]}: {T / 1e6:.2f} ms in {sum(cnt.values())} launches (serialised, cold caches: use the shares)")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"  {k:44s} {cnt[k]:5d} launches {v / T * 100:6.1f} %  {v / 1e6:9.3f} ms")
