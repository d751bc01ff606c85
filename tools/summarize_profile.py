#!/usr/bin/env python
"""Condenses `ncu --set full` reports (gpurun_out/*.ncu-rep) into small text summaries under profiles/."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def summarize(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{vals[i]} {units[i]}".strip()
        out.append(d)
    return out


def opmix(rep, top=16):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) < 3:
        return {}
    hdr, data = rows[1], rows[2:]
    ie, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
    tot, by = 0, {}
    for r in data:
        n = int(r[ie]); s = r[isrc].strip().split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        by[op] = by.get(op, 0) + n; tot += n
    return {"warp_instructions": tot, "share_pct": {k: round(100.0 * v / tot, 2) for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:top]}}


if __name__ == "__main__":
    for rep in sys.argv[1:]:
        print(json.dumps({"report": rep, "kernels": summarize(rep), "opcode_mix": opmix(rep)}, indent=1))
