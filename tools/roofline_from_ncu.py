#!/usr/bin/env python
"""Distil the numbers bench.py's `roofline` / `roofline_compute` objects quote from an .ncu-rep (ncu --set full
--clock-control none on bench.py) into profiles/roofline_pipes.json and profiles/roofline_traffic.json.

    python tools/roofline_from_ncu.py gpurun_out/prof.ncu-rep tensor-core "profiles/r02y_score_tc_bench_metrics.txt"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PICK = {
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "pipe_tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_math_pipe_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "registers_per_thread": "launch__registers_per_thread",
    "duration_us_under_ncu": "gpu__time_duration.sum",
    "sm_cycles_elapsed": "sm__cycles_elapsed.avg",
}


def num(s):
    return float(s.replace(",", ""))


def main():
    rep, key, source = sys.argv[1], sys.argv[2], sys.argv[3]
    match = {"tensor-core": "score_tc_kernel", "thread-per-query": "score_tq_kernel"}[key]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    hits = [r for r in rows[2:] if match in r[hdr.index("Kernel Name")]]
    if not hits:
        raise SystemExit(f"no {match} launch in {rep}")
    r = hits[-1]  # the last captured launch (warm)
    get = lambda m: num(r[hdr.index(m)])
    pipes = {k: get(m) for k, m in PICK.items() if m in hdr}
    if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond"):
        pipes["duration_us_under_ncu"] /= 1e3
    pipes["warp_instructions_per_launch"] = get("smsp__inst_executed.sum")
    pipes["kernel"] = r[hdr.index("Kernel Name")]
    pipes["source"] = source
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1)
    wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1)
    for name, entry in (("roofline_pipes.json", pipes),
                        ("roofline_traffic.json", {"kernel": pipes["kernel"], "dram_bytes_read": rd, "dram_bytes_write": wr,
                                                   "dram_bytes_per_launch": rd + wr, "source": source})):
        path = os.path.join(ROOT, "profiles", name)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[key] = entry
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)
        print(path, json.dumps(entry)[:300])


if __name__ == "__main__":
    main()
