#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/out.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    for r in rows[2:]:
        d = {hdr[i]: r[i] for i in idx}
        rd, wr = float(d.get("dram__bytes_read.sum", 0)), float(d.get("dram__bytes_write.sum", 0))
        print(d["Kernel Name"][:70], d["gpu__time_duration.sum"], units[hdr.index("gpu__time_duration.sum")],
              "dram r/w", rd, units[hdr.index("dram__bytes_read.sum")], wr, units[hdr.index("dram__bytes_write.sum")],
              "dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "lts%", d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main()
