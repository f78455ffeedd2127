#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table: one row per captured launch.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xyz.md]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__cycles_elapsed.max", "cyc elapsed"),
    ("sm__cycles_active.avg", "cyc active"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st math_thr"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st mio_thr"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st lg_thr"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st not_sel"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st dispatch"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st branch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st no_inst"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("| launch | kernel | " + " | ".join(k[1] for k in KEYS if k[0] in col) + " |")
    print("|---|---|" + "---|" * sum(1 for k in KEYS if k[0] in col))
    for n, r in enumerate(data):
        name = r[col["Kernel Name"]].split("(")[0][-40:]
        cells = []
        for k, _ in KEYS:
            if k in col:
                v, u = r[col[k]], units[col[k]]
                try:
                    f = float(v.replace(",", ""))
                    v = ("%.3g" % f) if abs(f) < 1e4 else ("%.4g" % f)
                except ValueError:
                    pass
                cells.append(v + (" " + u if u and u not in ("%", "inst", "cycle", "") else ""))
        print("| %d | %s | " % (n, name) + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
