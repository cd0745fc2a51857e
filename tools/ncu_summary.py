#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into a small text summary for profiles/.

    python tools/ncu_summary.py gpurun_out/r1/prof_c2.ncu-rep > profiles/r1_ncu_full_c2.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["wait", "short_scoreboard", "long_scoreboard", "not_selected", "no_instruction",
               "math_pipe_throttle", "mio_throttle", "branch_resolving", "barrier", "dispatch_stall",
               "lg_throttle", "membar", "sleeping"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (times are cold-cache, serialised launches)")
    for r in rows[2:]:
        print("\n== " + r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]:>16s} {units[i]}")
        stalls = []
        for s in STALL_NAMES:
            k = STALLS % s
            if k in hdr:
                stalls.append((float(r[hdr.index(k)] or 0), s))
        print("  warp stalls per issued instruction: " +
              ", ".join(f"{s}={v:.2f}" for v, s in sorted(stalls, reverse=True) if v >= 0.01))


if __name__ == "__main__":
    main()
