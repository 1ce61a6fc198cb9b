#!/usr/bin/env python
"""Turn the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv FIRST LAST out.csv "comment"
      per-kernel totals of launches FIRST..LAST (0-based, inclusive) of an `ncu --metrics gpu__time_duration.sum --csv` log
  python profiles/summarize.py full gpurun_out/prof.ncu-rep out.csv "comment"
      selected metrics of every kernel in an `ncu --set full` report (needs ncu on PATH; no GPU)
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_tma_ld.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(path, first, last, out, comment):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h, rows = rows[0], rows[1:]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    sel = rows[first:last + 1]
    tot = collections.OrderedDict()
    for r in sel:
        name = r[ki].split("(")[0].replace("cbs::", "").replace("void ", "")
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[vi].replace(",", "")) * 1e-6
    total = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write("# " + comment + "\n")
        f.write(f"# launches {first}..{last} of the process; sum of the serialised per-launch times: {total:.1f} ms\n")
        f.write("kernel,launches,total_ms,share_pct_of_serialised_sum,avg_ms\n")
        for k, (n, ms) in tot.items():
            f.write(f"{k},{n},{ms:.3f},{100 * ms / total:.1f},{ms / n:.4f}\n")


def full(rep, out, comment):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    ki = h.index("Kernel Name")
    with open(out, "w") as f:
        f.write("# " + comment + "\n")
        f.write("metric,unit," + ",".join(r[ki].split("(")[0].replace("cbs::", "").replace("void ", "") for r in data) + "\n")
        for m in KEEP:
            if m in h:
                i = h.index(m)
                f.write(f"{m},{units[i]}," + ",".join(r[i].replace(",", "") for r in data) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
