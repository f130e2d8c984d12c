#!/usr/bin/env python
"""Extract the metrics quoted in profiles/*/SUMMARY.md from .ncu-rep files: python scripts/ncu_summary.py name=file.ncu-rep ... > out.json"""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_shared_red.sum", "smsp__inst_executed_op_shared_ld.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
    "smsp__average_warp_latency_issue_stalled_membar.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_imc_miss.ratio", "smsp__average_warp_latency_issue_stalled_drain.ratio",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]

def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h == "Kernel Name":
            d[h] = {"value": v, "unit": u}
    return d

if __name__ == "__main__":
    res = {}
    for a in sys.argv[1:]:
        name, path = a.split("=", 1)
        res[name] = rep(path)
    json.dump(res, sys.stdout, indent=1)
