#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per-launch duration, DRAM bytes, L1/L2 hit rates,
occupancy, registers, top stall reasons.  Usage: ncu_summary.py file.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "launch__grid_size", "launch__block_size",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in data:
        lines.append(f"== {r[idx['Kernel Name']]}  grid {r[idx.get('Grid Size', 0)]} block {r[idx.get('Block Size', 0)]}")
        for k in KEYS:
            if k in idx:
                lines.append(f"   {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
        stalls = []
        for h, i in idx.items():
            if "issue_stalled_" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:6]:
            name = h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")
            lines.append(f"   stall cycles per issued instruction: {name:28s} {v:8.3f}")
        for k in ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
            if k in idx:
                lines.append(f"   {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")


if __name__ == "__main__":
    main()
