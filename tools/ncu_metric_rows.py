"""Print (kernel, metric, value) rows of an `ncu --csv --log-file` launch list, optionally filtered."""
import csv
import sys
rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for r in rows:
    if pat in r["Kernel Name"]:
        print(r["ID"], r["Kernel Name"].split("(")[0][-28:], r["Metric Name"], r["Metric Value"], r["Metric Unit"])
