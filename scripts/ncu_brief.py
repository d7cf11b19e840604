#!/usr/bin/env python3
"""Print the handful of counters we quote from an .ncu-rep:  python scripts/ncu_brief.py file.ncu-rep [more metrics]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum"]


def main():
    rep = sys.argv[1]
    want = WANT + sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    for r in rows[2:]:
        print(r[h.index("Kernel Name")][:100])
        for w in want:
            if w in h:
                print("   %-80s %s %s" % (w, r[h.index(w)], rows[1][h.index(w)]))


if __name__ == "__main__":
    main()
