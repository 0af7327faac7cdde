#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row.get("Metric Value", "0").replace(",", ""))
        except ValueError:
            continue
        unit = row.get("Metric Unit", "")
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row.get("Kernel Name", "")))
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':72s} {'n':>6s} {'total_us':>11s} {'share':>7s} {'avg_us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[:72]:72s} {v[0]:6d} {v[1]:11.1f} {v[1] / tot:7.3f} {v[1] / v[0]:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
