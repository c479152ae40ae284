#!/usr/bin/env python
"""Summarise an ncu launch-list csv (`--metrics gpu__time_duration.sum --csv --log-file x.csv`) as a markdown table:
share of device time, launch count and average duration per kernel.  Usage: launch_list_summary.py x.csv [skip_launches]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    kn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    unit = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[1:]:
        if len(r) <= mv or not r[idc].isdigit():
            continue
        v = float(r[mv].replace(",", ""))
        if r[unit] in ("us", "usecond"):
            v *= 1e3
        elif r[unit] in ("ms", "msecond"):
            v *= 1e6
        name = re.sub(r"\(.*", "", r[kn])[:90]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    print("| share | launches | avg ns | kernel |\n|---:|---:|---:|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 32]:
        print(f"| {100 * t / total:.1f}% | {n} | {t / n:.0f} | `{name}` |")
    print(f"\nTotal device time in the capture: {total / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches.")


if __name__ == "__main__":
    main()
