#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, on the CPU box) into the handful of numbers DESIGN.md / bench.py cite.
Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--sass N]"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for launch in rows[2:]:
        print("kernel:", launch[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:90s} {launch[i]:>16s} {units[i]}")
    if "--sass" in sys.argv:
        n = int(sys.argv[sys.argv.index("--sass") + 1])
        rows = page(rep, "source")
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = rows[2:]
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        c = Counter()
        for r in data:
            toks = r[ix["Source"]].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") else toks[0]
            c[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
        print("  opcode mix (warp instructions executed):")
        for k, v in c.most_common(n):
            print(f"    {k:10s} {v:14d} {100 * v / tot:5.1f}%")
        print("  hottest SASS by stall samples:")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
            print(f"    {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']][:80]}")


if __name__ == "__main__":
    main()
