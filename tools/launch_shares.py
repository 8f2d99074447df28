#!/usr/bin/env python
"""Kernel shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: tools/launch_shares.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t, n = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum" or int(r[col["ID"]]) < skip:
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        t[name] += v
        n[name] += 1
    total = sum(t.values())
    print(f"total {total:.2f} ms in {sum(n.values())} launches")
    for k in sorted(t, key=t.get, reverse=True):
        print(f"{t[k] / total * 100:6.1f} %  {t[k]:9.2f} ms  {n[k]:5d} x  {k}")


if __name__ == "__main__":
    main()
