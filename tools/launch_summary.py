#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share.
Usage: python tools/launch_summary.py gpurun_out/launches.csv [> profiles/xyz.md]"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    col = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            t = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[col["Metric Unit"]], 1.0)
        k = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.1f | %.2f | %.1f %% |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("\n(per-launch times are cold-cache and serialised by ncu: compare shares, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])
