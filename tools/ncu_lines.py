#!/usr/bin/env python
"""Per-source-line attribution from `ncu -i X --page source --csv --print-source cuda,sass -k regex:K` output:
warp instructions executed and stall samples summed over the SASS of each CUDA source line."""
import collections
import csv
import sys


def main(path, min_pct=0.7):
    agg = collections.OrderedDict()
    fname, hdr, last = "", None, ("", "")
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or len(r) < 8:
            continue
        line, src = r[0], r[1]
        if not line.strip():
            line, src = last
        last = (line, src)
        try:
            n = int(r[7] or 0)
            s = int(r[6] or 0)
        except ValueError:
            continue
        k = (fname, line)
        a = agg.setdefault(k, [0, 0, src.strip()])
        a[0] += n
        a[1] += s
        if src.strip():
            a[2] = src.strip()
    tot = sum(a[0] for a in agg.values()) or 1
    stot = sum(a[1] for a in agg.values()) or 1
    print("total warp insts %.2f M, samples %d" % (tot / 1e6, stot))
    for (f, l), (n, s, src) in agg.items():
        if 100.0 * n / tot >= min_pct or 100.0 * s / stot >= min_pct:
            print("%6.2fM %5.1f%% | smp %5.1f%% | %s:%s  %s" % (n / 1e6, 100.0 * n / tot, 100.0 * s / stot, f, l, src[:100]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.7)
