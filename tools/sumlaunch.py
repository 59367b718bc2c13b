#!/usr/bin/env python
"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/sumlaunch.py file.csv [last_n_launches]"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            u = r.get("Metric Unit", "ns")
            v = v / 1000.0 if u.startswith("ns") else (v if u.startswith("us") else v * 1000.0)
            rows.append((r["Kernel Name"].split("(")[0][:60], v))
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s n=%4d  %9.1f us  %5.1f%%" % (k, n, v, 100 * v / tot))
    print("total %.1f us over %d launches" % (tot, len(rows)))


if __name__ == "__main__":
    main()
