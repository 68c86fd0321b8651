#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
    python profiles/summarize.py report   gpurun_out/prof_x.ncu-rep          > profiles/rNN_x.md

`launches` aggregates a `--metrics gpu__time_duration.sum --csv` log per kernel (count, total, share).
`report` reads a `--set full` capture through `ncu -i ... --page raw --csv` and prints the metrics the
roofline discussion uses (duration, DRAM bytes, pipe utilisation, issue activity, stall reasons).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki]
        for cut in ("(const", "(T1", "(at::", "(long"):
            if cut in name:
                name = name[: name.index(cut)]
        agg.setdefault(name[:110], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print("| launches | total ms | share | avg us | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {len(v)} | {sum(v) / 1e6:.3f} | {100 * sum(v) / total:.1f}% | {sum(v) / len(v) / 1e3:.1f} | `{k}` |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ni]
        print(f"### `{name[:160]}`\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    print(f"| {h} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
