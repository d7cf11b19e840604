#!/usr/bin/env python3
"""profiles/r2_sweep.md from the sweep JSON files of round 2 (gpurun_out/r2_sweep_*.json, written by scripts/sweep.py)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def load(name):
    p = os.path.join(G, name)
    return json.load(open(p)) if os.path.isfile(p) else []


def table(rows, variants=None, title="", note=""):
    out = ["### " + title, ""]
    if note:
        out += [note, ""]
    out += ["| block | batch | variant | fwd ms | bwd ms (incl. weight-gradient GEMMs) | fwd+bwd algorithmic GB/s | % of 6550 |", "|---|---|---|---|---|---|---|"]
    for r in rows:
        if variants and r["variant"] not in variants:
            continue
        out.append("| %dx%d^2 | %d | %s | %.3f | %.3f | %.0f | %.1f |" % (
            r["c"], r["h"], r["n"], r["variant"], r["fwd_ms"], r["bwd_ms"], r["fwd_bwd_gbs"], 100 * r["frac_peak"]))
    return out + [""]


def main():
    doc = ["# Round 2 sweeps (one B200, `scripts/sweep.py`, CUDA events, L2 flushed before every call, median of 8)", "",
           "`old` = round-1 paths (tunable `tile_kind=2`): cluster kernels for 128x28^2 / 256x14^2, streaming plane kernels + batched GEMMs",
           "for 512x7^2. `auto` = shipped selection. Variants starting with another letter force the tile pipeline with the tunables",
           "named in `scripts/sweep.py`. The backward column includes the weight-gradient GEMM launch that follows the block's kernel.", ""]
    doc += table(load("r2_sweep_final.json"), title="Final selection at the end of round 2 (`auto`) vs cluster / streaming kernels only (`old` = `tile_kind=2`) vs pipeline forced (`sw_tile`)",
                 note="After the second half of round 2 (`profiles/r2_cluster_kernels.md`): cluster kernels with the FC weights in shared memory and the "
                      "st.async exchange, weight-gradient GEMMs inside the backward pipeline launch.")
    doc += table(load("r2_sweep_pol2.json"), title="Mid-round selection vs round-1 paths (before the cluster-kernel work)")
    doc += table(load("r2_sweep_f.json"), title="Forward chunk size (f_c28 / f_c56 / f_c100 = 28 / 56 / 100 KB items), tile pipeline forced",
                 note="56 KB wins from ~190 MB per modality, 28 KB below; `old` rows show where the pipeline starts to pay.")
    doc += table(load("r2_sweep_t.json") + load("r2_sweep_c.json") + load("r2_sweep_s.json"),
                 variants={"x_ronly", "x_sonly", "q_rload", "q_rload_c56", "q_rload_c14", "c_rload_x2", "c_rload_x4", "c_rload_x7",
                           "c_rload_c56_x2", "s_rload_slots4", "s_rload_slots2"},
                 title="Stage rates in isolation (measurement modes, results are garbage)",
                 note="`x_ronly` / `x_sonly`: R stage / S stage alone (no GEMM work, no dependencies). `q_rload*`: R-stage loads only, "
                      "workers idle: the rate follows the ITEM size (14 / 28 / 56 KB), not the number of bulk copies per item (`c_rload_x2/4/7`, "
                      "`c_rload_c56_x2`) nor the ring depth (`s_rload_slots*`): ~0.85 us of loader time per item. The GB/s columns are "
                      "computed for a full fwd/bwd and do not apply here; compare the ms. R stage alone at 128x28^2 reads 822 MB: "
                      "0.164 ms = 5.0 TB/s with 28 KB items, 0.139 ms = 5.9 TB/s with 56 KB.")
    doc += table(load("r2_sweep_sw.json"), title="GEMM CTAs joining the stream role; number of GEMM CTAs; lag",
                 note="`noswitch` = GEMM CTAs exit when their tickets run out (round-2 state before the change).")
    doc += table(load("r2_sweep_k.json"), title="Split-K inside the pipeline, tile size (all slower than the default)")
    doc += table(load("r2_sweep_d.json"), title="Tickets per draw (d_drawN), chunk size for both directions (d_c40 / d_c56)")
    doc += table(load("r2_sweep_p.json"), variants={"p_tile_p1", "p_tile_p2", "sw_tile"},
                 title="L2 eviction policy of the R-stage reads (p1 evict_first, p2 evict_normal, default evict_last): neutral")
    doc += table(load("r2_sweep_x.json"), title="Dependency / store experiments at 512x7^2 (x_nodeps1 no S dependency, 2/3 idle workers, 4 no stores)",
                 note="Even without dependencies the launch cannot end before the last tile's FC chain: the block is chain-latency bound.")
    open(os.path.join(ROOT, "profiles", "r2_sweep.md"), "w").write("\n".join(doc) + "\n")
    print("wrote profiles/r2_sweep.md, %d lines" % len(doc))


if __name__ == "__main__":
    main()
