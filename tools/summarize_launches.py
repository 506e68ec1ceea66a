#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, device time, share.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {n - skip} launches after skipping {skip}, {tot:.1f} us total device time (cold-cache, serialised)")
    print(f"# {'us':>10} {'launches':>8} {'share':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {v[0]:8d} {100 * v[1] / tot:5.1f}%  {k[:120]}")


if __name__ == "__main__":
    main()
