#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid):
python profiles/launch_summary.py gpurun_out/launches.csv > profiles/launches_rNN.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name if len(name) <= 90 else name[:87] + "..."


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        try:
            ns = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        acc.setdefault((short(r[ki]), r[gi]), []).append(ns)
    total = sum(sum(v) for v in acc.values())
    print("| kernel | grid | launches | avg us | total us | share |")
    print("|---|---|---|---|---|---|")
    for (k, g), v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {g} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / total:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])
