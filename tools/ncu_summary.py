#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into the text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_score_tq   # writes <prefix>_metrics.txt, <prefix>_hot_sass.txt
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
           "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    with open(prefix + "_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {rep}\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n== {name}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:95s} {units[i]:16s} {r[i]}\n")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    for n, r in enumerate(body):  # several captured launches: keep the first kernel's block only
        if r and r[0] == "Kernel Name":
            body = body[:n]
            break
    data = [r for r in body if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    with open(prefix + "_hot_sass.txt", "w") as f:
        f.write(f"# {rows[0][1] if len(rows[0]) > 1 else ''}\n# warp-state samples: {tot}; instructions with >= 0.25% of the samples\n")
        f.write(f"# {'addr':>6} {'samples':>7} {'executed':>9}  stalls(>3)  SASS\n")
        for r in data:
            s = int(r[ix["# Samples"]] or 0)
            if tot and s >= 0.0025 * tot:
                st = {h[6:]: int(r[ix[h]] or 0) for h in stalls if int(r[ix[h]] or 0) > 3}
                f.write(f"{r[ix['Address']][-5:]:>8} {s:7d} {int(r[ix['Instructions Executed']] or 0):9d}  {st}  {r[ix['Source']]}\n")


if __name__ == "__main__":
    main()
