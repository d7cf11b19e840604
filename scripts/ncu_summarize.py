#!/usr/bin/env python3
"""Turn .ncu-rep captures into the text summaries kept under profiles/ and update profiles/traffic.json.
   python scripts/ncu_summarize.py blocks    # gpurun_out/r2_block_{128,256,512}.ncu-rep -> profiles/r2_ncu_blocks.md + traffic.json
   python scripts/ncu_summarize.py launches  # gpurun_out/r2_ncu_launches.csv -> profiles/r2_ncu_launch_list_summary.txt"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
     "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
     "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"name": r[h.index("Kernel Name")]}
        for m in M:
            hm = [x for x in h if x == m or x.endswith("." + m)]
            hm = [x for x in hm if r[h.index(x)] not in ("", "no data", "n/a")]
            if hm:
                v = float(r[h.index(hm[0])].replace(",", ""))
                d[m] = v * UNIT.get(units[h.index(hm[0])], 1.0)
        res.append(d)
    return res


def blocks():
    n = 1024
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    db = json.load(open(tfile)) if os.path.isfile(tfile) else {}
    lines = ["# Round 2: `ncu --set full` of the dominant kernel of every block, both directions, batch 1024",
             "", "Command: `bash scripts/ncu_r2.sh` (one `scripts/tile_one.py --path auto --c C --h H --n 1024 --bwd --reps 1` per shape,",
             "default cache control = caches flushed before every replay, `--clock-control none`).  Reports: `gpurun_out/r2_block_<C>.ncu-rep`",
             "(scratch); this table is the committed summary.  u = N*C*HW*4 bytes of one modality; algorithmic bytes fwd 4u, bwd 6u.", "",
             "| block | dir | kernel | grid x block | time (us) | DRAM read (MB) | DRAM write (MB) | read+write / algorithmic | real DRAM TB/s | algorithmic TB/s | DRAM %peak (ncu) | L2 hit % | tensor pipe active % |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for c, h in ((128, 28), (256, 14), (512, 7)):
        rep = os.path.join(ROOT, "gpurun_out", "r2_block_%d.ncu-rep" % c)
        ks = raw(rep)
        assert len(ks) == 2, [k["name"] for k in ks]
        u = n * c * h * h * 4
        for d_, k, units in (("fwd", ks[0], 4), ("bwd", ks[1], 6)):
            tr = k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]
            t = k["gpu__time_duration.sum"]
            short = k["name"].split("(")[0].split("::")[-1]
            lines.append("| %dx%d^2 | %s | `%s` | %d x %d | %.1f | %.1f | %.1f | %.3f | %.2f | %.2f | %.1f | %.1f | %.2f |" % (
                c, h, d_, short, k["launch__grid_size"], k["launch__block_size"], t * 1e6, k["dram__bytes_read.sum"] / 1e6,
                k["dram__bytes_write.sum"] / 1e6, tr / (units * u), tr / t / 1e12, units * u / t / 1e12,
                k.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0), k.get("lts__t_sector_hit_rate.pct", 0),
                k.get("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 0)))
            db["%dx%d^2/fused_%s@%d" % (c, h, d_, n)] = {
                "traffic_bytes_per_launch": tr, "dram_read_bytes": k["dram__bytes_read.sum"],
                "dram_write_bytes": k["dram__bytes_write.sum"], "algorithmic_bytes": units * u, "ncu_time_us": t * 1e6,
                "kernel": short, "source": "profiles/r2_ncu_blocks.md (ncu --set full, caches flushed per replay, scripts/ncu_r2.sh)"}
    open(os.path.join(ROOT, "profiles", "r2_ncu_blocks.md"), "w").write("\n".join(lines) + "\n")
    json.dump(db, open(tfile, "w"), indent=1)
    print("\n".join(lines))


def launches():
    path = os.path.join(ROOT, "gpurun_out", "r2_ncu_launches.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[h.index("Kernel Name")].split("(")[0].split("::")[-1]
        v = float(r[h.index("Metric Value")].replace(",", ""))
        unit = r[h.index("Metric Unit")]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 3 --no-train --no-stats",
           "(first 600 launches of the process: warm-up, the timed graph replays, the batch sweep, per-block profiling; per-launch times are",
           " cold-cache and serialised -- the SHARE is what must agree with bench.py's per-block event times)", "",
           "%-60s %8s %12s %8s" % ("kernel", "launches", "total us", "share")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-60s %8d %12.1f %7.1f%%" % (k[:60], a[0], a[1], 100 * a[1] / tot))
    open(os.path.join(ROOT, "profiles", "r2_ncu_launch_list_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    {"blocks": blocks, "launches": launches}[sys.argv[1]]()
