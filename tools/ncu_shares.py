#!/usr/bin/env python3
"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    tools/ncu_shares.py <launches.csv> [top-n]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ci = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[ci["Kernel Name"]].split("(")[0][:72]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        u = r[ci["Metric Unit"]]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[u]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        print("%-74s n=%5d  %10.3f ms  %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))


main()
