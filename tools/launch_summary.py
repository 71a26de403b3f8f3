#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.OrderedDict()
T = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1
    tot[name][1] += v
    T += v
print("total %.3f ms over %d launches" % (T, sum(c for c, _ in tot.values())))
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-34s x%-4d %9.3f ms %5.1f%%" % (k, c, v, 100 * v / T))
