#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the captured window)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hdr_i]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr_i + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} n={v[0]:5d} us={v[1]:10.1f} avg_us={v[1] / v[0]:8.1f} share={100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
