#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few metrics the roofline
discussion uses; prints one JSON object per kernel launch.  Usage: ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "dram__bytes_read.sum.per_second": "dram_read_rate", "dram__bytes_write.sum.per_second": "dram_write_rate",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_sectors_op_atom.sum": "l2_atom_sectors", "lts__t_sectors_op_red.sum": "l2_red_sectors",
    "lts__t_sectors_op_read.sum": "l2_read_sectors", "lts__t_sectors_op_write.sum": "l2_write_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum": "local_ld_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum": "local_st_sectors",
    "smsp__inst_executed.sum": "warp_insts",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_inst",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
}


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"file": path.split("/")[-1]}
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d["kernel"] = v
            if h in KEYS:
                try:
                    d[KEYS[h]] = f"{float(v):.6g} {u}".strip()
                except ValueError:
                    d[KEYS[h]] = v
        res.append(d)
    return res


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for d in summarize(p):
            print(json.dumps(d))
