#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (stdin or file).

    ncu -i gpurun_out/prof.ncu-rep --page source --csv | python profiles/hotspots.py [N]
"""
import csv
import sys


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    rows = list(csv.reader(sys.stdin))
    kernel, hdr = None, None
    body = []
    for r in rows:
        if r and r[0] == "Kernel Name":
            if body:
                report(kernel, hdr, body, n)
            kernel, hdr, body = r[1], None, []
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) >= len(hdr) - 2:
            body.append(r)
    if body:
        report(kernel, hdr, body, n)


def report(kernel, hdr, body, n):
    ci, si, ei = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    f = lambda x: float(x) if x not in ("", None) else 0.0
    tot = sum(f(r[ci]) for r in body) or 1.0
    print("== %s\n   %d SASS instructions, %d samples, %.3g warp-instructions executed" % (
        kernel[:110], len(body), tot, sum(f(r[ei]) for r in body)))
    agg = {}
    for r in body:
        for i in stalls:
            agg[hdr[i]] = agg.get(hdr[i], 0.0) + f(r[i])
    print("   stall mix: " + ", ".join("%s %.0f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]))
    for idx, r in sorted(enumerate(body), key=lambda x: -f(x[1][ci]))[:n]:
        top = max(stalls, key=lambda i: f(r[i]))
        print("%5.1f%%  #%-5d %-14s %s" % (100 * f(r[ci]) / tot, idx, hdr[top][6:], r[si][:100]))


if __name__ == "__main__":
    main()
