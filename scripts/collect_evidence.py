#!/usr/bin/env python3
"""Turn gpurun_out/ of scripts/gpu_final.sh into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

d = json.loads(open(os.path.join(G, "bench.json")).read().splitlines()[-1])
json.dump(d, open(os.path.join(P, "%s_bench_b200_n1.json" % tag), "w"), indent=1)
r = json.loads(open(os.path.join(G, "bench_ref.json")).read().splitlines()[-1])
json.dump(r, open(os.path.join(P, "%s_bench_reference_arm.json" % tag), "w"), indent=1)
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "frac_of_measured_hbm_peak", "clocks")})
print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "ref arm", r["value"])
print({k: v for k, v in d["roofline"].items() if k != "note"})
print("sweep", {k: (round(v["value"]), round(v["frac_of_measured_hbm_peak"], 3)) for k, v in d["sweep"].items()})
print("train", {k: d["train"][k] for k in ("samples_per_s", "ms_per_step", "speedup_vs_cpu_reference")})

rows = json.load(open(os.path.join(G, "sweep.json")))
with open(os.path.join(P, "%s_sweep.md" % tag), "w") as f:
    f.write("# MMTM fwd+bwd sweep on one B200 (BASELINE configs[1]): C-ABI calls, CUDA events per call, L2 flushed "
            "between iterations\n\n`auto` = shipped path selection (cluster/L2-resident kernels for 128x28^2 and "
            "256x14^2, streaming for 512x7^2); `stream_nochunk` = two-pass streaming kernels forced.\nalgorithmic GB/s "
            "= 4u (fwd), 6u (bwd), 10u (fwd+bwd) per wall time of the whole call (kernels + FC GEMMs + column sums); "
            "peak = 6550.1 GB/s measured copy.\n\n| C x HW | N | path | fwd ms | fwd GB/s | bwd ms | bwd GB/s | "
            "fwd+bwd GB/s | % of peak |\n|---|---|---|---|---|---|---|---|---|\n")
    for x in rows:
        f.write("| %dx%d^2 | %d | %s | %.3f | %.0f | %.3f | %.0f | %.0f | %.1f |\n" % (
            x["c"], x["h"], x["n"], x["variant"], x["fwd_ms"], x["fwd_gbs"], x["bwd_ms"], x["bwd_gbs"],
            x["fwd_bwd_gbs"], 100 * x["frac_peak"]))

# launch list
lrows = [x for x in csv.reader(open(os.path.join(G, "launches.csv"))) if len(x) > 10]
hdr = lrows[0]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for x in lrows[1:]:
    short = x[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("gml::<unnamed>::", "").replace("gml::", "")
    v = float(x[idx["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}[x[idx["Metric Unit"]]]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += v
mine = {k: v for k, v in agg.items() if any(t in k for t in ("l2_", "fused_", "plane_", "quad_", "gemm", "colsum",
                                                               "sqnorm", "fill", "running"))}
tot = sum(v[1] for v in mine.values())
with open(os.path.join(P, "%s_ncu_launch_list_summary.txt" % tag), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --no-train --steps 2 "
            "--warmup 3\n(first 400 launches of the process: eager warm-up + graph capture + replays at batch 1024, then "
            "the batch-32/256 sweep; per-launch times are cold-cache and serialised -> compare SHARES)\n\nlibrary kernels "
            "only, share of their summed device time:\n")
    for k, v in sorted(mine.items(), key=lambda kv: -kv[1][1]):
        f.write("%6d launches %10.1f us %5.1f%%  %s\n" % (v[0], v[1], 100 * v[1] / tot, k[:110]))
subprocess.run(["cp", os.path.join(G, "launches.csv"), os.path.join(P, "%s_ncu_launches.csv" % tag)])
print(open(os.path.join(P, "%s_ncu_launch_list_summary.txt" % tag)).read())

# dominant kernel
raw = subprocess.run(["ncu", "-i", os.path.join(G, "dominant_prof.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__cluster_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
tob = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
traffic = []
with open(os.path.join(P, "%s_ncu_dominant_l2_bwd_128x28_n1024.txt" % tag), "w") as f:
    f.write("ncu --set full --clock-control none --cache-control none --import-source on -k regex:l2_bwd -c 2 : python "
            "scripts/sweep.py --batches 1024 --iters 1 --variants auto\ndominant kernel of bench.py's step: l2_bwd_kernel "
            "(cluster of 4, L2-resident grad_out), 128x28^2, N=1024; algorithmic bytes per launch 6u = 2466.25 MB\n")
    for x in rr[2:]:
        f.write("----\n")
        for w in want:
            if w in idx:
                f.write("  %-62s %s %s\n" % (w, x[idx[w]][:90], units[idx[w]]))
        st = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
        vals = sorted([(float(x[idx[h]].replace(",", "")) if x[idx[h]] else 0, h) for h in st], reverse=True)[:6]
        f.write("  top stall reasons (warps per issue): " + ", ".join(
            "%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
            for v, h in vals) + "\n")
        t = tob(x[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + tob(
            x[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic.append(t)
        f.write("  => HBM traffic per launch %.1f MB = %.3f x algorithmic\n" % (t / 1e6, t / 2466250752))
json.dump({"128x28^2/fused_bwd@1024": {"traffic_bytes_per_launch": sum(traffic) / len(traffic),
                                        "source": "profiles/%s_ncu_dominant_l2_bwd_128x28_n1024.txt (ncu --set full, "
                                                  "--cache-control none)" % tag}},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "%s_ncu_dominant_l2_bwd_128x28_n1024.txt" % tag)).read()[-900:])

# tcgen05 FC GEMM
rep = os.path.join(G, "umma_gemm_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "gpu__time_duration.sum",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "greedy_multimodal_learning_b200", "csrc", "build",
                                                              "gemm_kernels.o")], capture_output=True, text=True).stdout
    counts = collections.Counter()
    for line in sass.splitlines():
        for m in ("UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "LDGSTS", "HMMA.1688.F32.TF32", "SYNCS", "UCGABAR"):
            if m in line:
                counts[m] += 1
    with open(os.path.join(P, "%s_ncu_umma_fc_gemm.txt" % tag), "w") as f:
        f.write("ncu --set full --clock-control none --cache-control none --import-source on -k regex:gemm_umma -c 2 : "
                "python scripts/gemm_accuracy.py\ntcgen05 3xTF32 FC GEMM (gemm_umma_kernel), C[1024,512] = A[1024,1024] "
                "B[512,1024]^T, 32 tiles x 4-way split-K as 4-CTA clusters\n")
        for x in rr[2:]:
            f.write("----\n")
            for w in want:
                if w in idx:
                    f.write("  %-62s %s %s\n" % (w, x[idx[w]][:90], units[idx[w]]))
        f.write("\nSASS mnemonics in gemm_kernels.o (cuobjdump -sass): " +
                ", ".join("%s x%d" % kv for kv in sorted(counts.items())) + "\n")
        acc = os.path.join(G, "gemm_accuracy.log")
        if os.path.exists(acc):
            f.write("\nscripts/gemm_accuracy.py on the same box (max abs error vs float64 on N(0,1) operands; time per "
                    "gml_fc_gemm call incl. ~11 us of launch floor; effective fp32 TFLOP/s):\n" + open(acc).read())
    print(open(os.path.join(P, "%s_ncu_umma_fc_gemm.txt" % tag)).read())
