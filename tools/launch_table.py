#!/usr/bin/env python
"""Prints the launches of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list
(steps are delimited by stft_mel_kernel launches) with durations, plus a per-kernel aggregate."""
import collections
import csv
import re
import sys


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return re.sub(r"\(.*", "", n)


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [short(x["Kernel Name"]) for x in rows]
    durs = [float(x["Metric Value"]) / 1e3 for x in rows]
    idx = [i for i, n in enumerate(names) if n.startswith("stft_mel")]
    s, e = idx[-1], len(rows)
    tot = sum(durs[s:e])
    print("launches in step: %d, serialised total %.1f us" % (e - s, tot))
    agg = collections.OrderedDict()
    for i in range(s, e):
        print("%8.1f us  stream %s  %s %s %s" % (durs[i], rows[i]["Stream"], names[i], rows[i]["Grid Size"], rows[i]["Block Size"]))
        a = agg.setdefault(names[i], [0, 0.0])
        a[0] += 1
        a[1] += durs[i]
    print("---- aggregate ----")
    for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%8.1f us %3dx %5.1f%%  %s" % (d, c, 100 * d / tot, k))


if __name__ == "__main__":
    main(sys.argv[1])
