#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.1f us total device time (cold-cache, serialised: compare shares)" % (path, len(rows) - 1, tot))
    print("%-72s %5s %11s %9s %6s" % ("kernel", "n", "total_us", "mean_us", "share"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %5d %11.1f %9.1f %6.3f" % (n, a[0], a[1], a[1] / a[0], a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
