#!/usr/bin/env python
"""Summarise one .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"== {rep}: {d.get('Kernel Name', '?')[:90]}")
            for k in KEYS:
                if k in d:
                    print(f"  {k:90s} {d[k]:>16s} {u[k]}")
            st = sorted(((float(v), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
            print("  stalls (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))


if __name__ == "__main__":
    main()
