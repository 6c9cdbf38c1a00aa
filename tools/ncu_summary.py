#!/usr/bin/env python
"""Text summary of one kernel of an .ncu-rep (run here, no GPU needed): the metrics DESIGN.md quotes.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "header line" > profiles/x.ncu.txt"""
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "gpu__time_duration.sum",
    "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def main():
    rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    if header:
        print(header + "\n")
    for k in KEYS:
        if k in col:
            print(f"{k:<88}{vals[col[k]]:>24} {units[col[k]]}")
    print("-- warp stall reasons (avg warps per issue-active cycle) --")
    for h in sorted(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            print(f"{h:<88}{vals[col[h]]:>24}")


if __name__ == "__main__":
    main()
