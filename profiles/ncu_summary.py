#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernel:", [r[hdr.index("Kernel Name")] for r in data])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {units[i]:16s} {[r[i] for r in data]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            vals = [float(r[i]) for r in data]
            if max(vals) >= 0.05:
                print(f"{h:75s} {[round(v, 3) for v in vals]}")


if __name__ == "__main__":
    main(sys.argv[1])
