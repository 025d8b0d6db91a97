#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into launches / total ms / share per kernel.
usage: python tools/launch_summary.py gpurun_out/x.csv ["header line"] > profiles/x_summary.txt"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) >= 15 and r[12] == "gpu__time_duration.sum"]
acc = OrderedDict()
for r in rows:
    ns = float(r[14].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[13], 1.0)
    k = acc.setdefault(r[4], [0, 0.0]); k[0] += 1; k[1] += ns
tot = sum(v[1] for v in acc.values())
if len(sys.argv) > 2: print(sys.argv[2])
for name, (n, ns) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-62s launches %4d total %10.3f ms share %5.1f%%" % (name[:60], n, ns / 1e6, 100.0 * ns / tot))
