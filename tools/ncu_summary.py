#!/usr/bin/env python3
"""Summarise an .ncu-rep (from `ncu --set full`) into a small CSV/markdown for profiles/.
usage: ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_full_rNN.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct"] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in
        ("long_scoreboard", "short_scoreboard", "barrier", "no_instruction", "math_pipe_throttle", "mio_throttle", "wait", "not_selected")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    lines = ["# ncu --set full summary of %s" % rep, "", "| kernel | " + " | ".join(k.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active", "") for k in KEYS) + " |",
             "|---|" + "---|" * len(KEYS)]
    for r in rows[2:]:
        if not r or name_i is None:
            continue
        vals = []
        for k in KEYS:
            i = idx.get(k)
            vals.append((r[i] + (" " + units[i] if units[i] else "")) if i is not None and i < len(r) else "n/a")
        lines.append("| %s | %s |" % (r[name_i][:60], " | ".join(vals)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
