#!/usr/bin/env python
"""Opcode histogram (warp instructions executed, stall samples) per kernel from `ncu --page source --csv` output.
usage: ncu_sass_hist.py src.csv [kernel-substring] [instance]"""
import collections
import csv
import sys


def kernels(path):
    out, cur = [], None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    return out


def main():
    ks = kernels(sys.argv[1])
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ks = [k for k in ks if sub in k["name"]]
    k = ks[inst]
    h = {n: i for i, n in enumerate(k["hdr"])}
    ops = collections.Counter()
    samples = collections.Counter()
    tot = 0
    for r in k["rows"]:
        sass = r[h["Source"]].strip()
        toks = sass.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.rstrip(";")
        n = int(r[h["Instructions Executed"]] or 0)
        ops[op] += n
        samples[op] += int(r[h["# Samples"]] or 0)
        tot += n
    print(k["name"][:90], " total warp insts %.2f M" % (tot / 1e6), " SASS lines", len(k["rows"]))
    stot = sum(samples.values()) or 1
    for op, n in ops.most_common(45):
        print("%-28s %9.3f M %5.1f%%   samples %5.1f%%" % (op, n / 1e6, 100.0 * n / tot, 100.0 * samples[op] / stot))


if __name__ == "__main__":
    main()
