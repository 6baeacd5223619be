#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU): one block of key counters per profiled launch.
    python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s: %d profiled launches (ncu --set full --clock-control none; per-launch, cold cache)" % (path, len(rows) - 2))
    for r in rows[2:]:
        print("kernel: %s" % r[idx["Kernel Name"]])
        for w in WANT:
            if w in idx:
                print("  %-95s %s %s" % (w, r[idx[w]], units[idx[w]]))
        try:
            rd, wr = float(r[idx["dram__bytes_read.sum"]].replace(",", "")), float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            print("  traffic = dram read + write = %.1f %s" % (rd + wr, units[idx["dram__bytes_read.sum"]]))
        except Exception:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
