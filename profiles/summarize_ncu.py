#!/usr/bin/env python
"""Summarise an ncu report (read here on the CPU box) into the handful of numbers quoted in
DESIGN.md / bench.py:  python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {path}: {len(data)} captured launches (ncu --set full --clock-control none)")
    for r in data:
        print(f"\n== {r[name_i]}")
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = (r[i], units[i])
                print(f"{w:88s} {r[i]:>18s} {units[i]}")
        if "dram__bytes_read.sum" in vals:
            rd = to_bytes(*vals["dram__bytes_read.sum"])
            wr = to_bytes(*vals["dram__bytes_write.sum"])
            t_us = float(vals["gpu__time_duration.sum"][0])
            t_s = t_us * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(vals["gpu__time_duration.sum"][1], 1e-6)
            print(f"{'traffic = dram read + write':88s} {rd + wr:18.0f} byte")
            print(f"{'dram GB/s over the launch':88s} {(rd + wr) / t_s / 1e9:18.1f} GB/s")


if __name__ == "__main__":
    main(sys.argv[1])
