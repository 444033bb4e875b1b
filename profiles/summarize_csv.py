#!/usr/bin/env python
"""Table of the metrics quoted in DESIGN.md from `ncu -i report.ncu-rep --page raw --csv` output
(the .ncu-rep files themselves are too large to bring back from the GPU box).
Usage: python profiles/summarize_csv.py raw.csv [columns] > profiles/<name>.txt"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64inst%"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_no_inst"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
        ("smsp__inst_executed_op_local_ld.sum", "local_ld"), ("smsp__inst_executed_op_local_st.sum", "local_st"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn")]


def main(path, columns=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"kernel: {r[ik]}")
        for m, short in WANT:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {short:14s} {r[i]:>18s} {units[i]:10s} {m}")
        if columns:
            sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            b = float(r[ir]) * sc[units[ir]] + float(r[iw]) * sc[units[iw]]  # (the two columns may carry different units)
            print(f"  dram bytes per (column, layer) at {columns} columns x 16 layers: {b / (int(columns) * 16):.0f}")


if __name__ == "__main__":
    main(*sys.argv[1:])
