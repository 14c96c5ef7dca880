#!/usr/bin/env python
"""Table of one `ncu --set full` launch per kernel from the raw-page CSV (ncu -i x.ncu-rep --page raw --csv):
python profiles/ncu_table.py profiles/ncu_kernels_r01_raw.csv"""
import csv
import re
import sys

COLS = [("us", "gpu__time_duration.sum", 1.0), ("rd", "dram__bytes_read.sum", None), ("wr", "dram__bytes_write.sum", None),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("regs", "launch__registers_per_thread", 1.0),
        ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1.0), ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
        ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1.0), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0)]
STALLS = [("long_sb", "long_scoreboard"), ("math", "math_pipe_throttle"), ("wait", "wait"), ("short_sb", "short_scoreboard"), ("lg_thr", "lg_throttle"),
          ("not_sel", "not_selected")]


def num(x):
    return float(x.replace(",", ""))


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " | " + " | ".join(s[0] for s in STALLS) + " |")
    print("|" + "---|" * (1 + len(COLS) + len(STALLS)))
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("ig::", "")
        cells = []
        for label, key, _ in COLS:
            v, u = num(r[ix[key]]), units[ix[key]]
            if label in ("rd", "wr"):
                scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}[u]
                cells.append(f"{v * scale:.0f}M")
            elif label == "us":
                cells.append(f"{v * {'us': 1.0, 'ms': 1e3, 'ns': 1e-3}[u]:.1f}")
            else:
                cells.append(f"{v:.1f}")
        for _, key in STALLS:
            cells.append(f"{num(r[ix[f'smsp__average_warps_issue_stalled_{key}_per_issue_active.ratio']]):.1f}")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
