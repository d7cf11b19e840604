#!/usr/bin/env python3
"""Per-CTA cycle breakdown of the tile pipeline (debug tunable "tile_stats_ptr"): where stream CTAs (loader: slot
wait / dependency wait; workers: data wait / reduce / scale) and GEMM CTAs (dependency wait / main loop / epilogue)
spend their time.   python scripts/tile_stats.py [--n 1024] [--tunables tile_lag=3,tile_gemm_ctas=24]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import SHAPES, BlockBuffers  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--tunables", default="")
    args = ap.parse_args()
    lib = L.load()
    dev = torch.device("cuda:0")
    for kv in filter(None, args.tunables.split(",")):
        k, v = kv.split("=")
        L.check(lib.gml_set_tunable(k.encode(), int(v)))
    stats = torch.zeros(1024 * 16, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for c, h in SHAPES:
        b = BlockBuffers(torch, L, args.n, c, h, dev, seed=c)
        P = lambda t: t.data_ptr()
        w, dw = b.w, b.dw

        def fwd():
            L.check(lib.gml_mmtm_fwd(P(b.a), P(b.b), P(b.a_out), P(b.b_out), P(w[0]), P(w[1]), P(w[2]), P(w[3]), P(w[4]),
                                     P(w[5]), P(b.z), P(b.hid), P(b.g_a), P(b.g_b), P(b.gate_sum), P(b.run_v), P(b.run_s),
                                     0, None, None, P(b.fws), b.fws_bytes, b.dims, 0, 1.0, L.F_FORCE_TILE, st))

        def bwd():
            L.check(lib.gml_mmtm_bwd(P(b.go_a), P(b.go_b), P(b.a), P(b.b), P(w[0]), P(w[2]), P(w[4]), P(b.z), P(b.hid),
                                     P(b.g_a), P(b.g_b), None, None, None, None, P(b.d_a), P(b.d_b), P(dw[0]), P(dw[1]),
                                     P(dw[2]), P(dw[3]), P(dw[4]), P(dw[5]), P(b.ws), b.ws_bytes, b.dims, 0, 1.0,
                                     L.F_FORCE_TILE, st))

        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            fn(); fn()
            torch.cuda.synchronize()
            stats.zero_()
            L.check(lib.gml_set_tunable(b"tile_stats_ptr", stats.data_ptr()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            L.check(lib.gml_set_tunable(b"tile_stats_ptr", 0))
            s_all = stats.view(-1, 16).cpu().double()
            nsm = torch.cuda.get_device_properties(0).multi_processor_count
            s, dbg = s_all[:nsm], s_all[nsm:2 * nsm]
            dbg = dbg[(s[:, 7] > 0) & (s[:, 0] == 1)]
            s = s[s[:, 7] > 0]
            g, r = s[s[:, 0] == 1], s[s[:, 0] == 0]
            ms = e0.elapsed_time(e1)
            mhz = 1.0  # report in kilo-cycles
            print("C=%d H=%d N=%d %s: %.3f ms (incl. weight-gradient GEMMs in bwd)" % (c, h, args.n, name, ms))
            if len(r):
                print("  stream CTAs %3d: total %7.0f kcyc | items %5.0f | loader slot-wait %6.0f dep-wait %6.0f | "
                      "warp0 data-wait %6.0f reduce %6.0f scale %6.0f" % (
                          len(r), r[:, 7].mean() / 1e3, r[:, 1].mean(), r[:, 2].mean() / 1e3, r[:, 3].mean() / 1e3,
                          r[:, 4].mean() / 1e3, r[:, 5].mean() / 1e3, r[:, 6].mean() / 1e3))
            if len(g):
                print("  GEMM   CTAs %3d: total %7.0f kcyc | items %5.1f | dep-wait %6.0f main %6.0f epilogue %6.0f "
                      "(kcyc per item: main %.1f epi %.1f)" % (
                          len(g), g[:, 7].mean() / 1e3, g[:, 1].mean(), g[:, 2].mean() / 1e3, g[:, 3].mean() / 1e3,
                          g[:, 4].mean() / 1e3, g[:, 3].sum() / max(g[:, 1].sum(), 1) / 1e3,
                          g[:, 4].sum() / max(g[:, 1].sum(), 1) / 1e3))
                print("        epilogue split: partial+fold %6.0f store %6.0f fence %6.0f sync+signal %6.0f kcyc" % tuple(
                    (g[:, 8 + i].mean() / 1e3) for i in range(4)))
                if len(dbg):
                    print("        main loop, control thread: wait-full %6.0f mma-issue %6.0f tma-issue(+empty wait) %6.0f | "
                          "worker0: wait-raw %6.0f split %6.0f drain %6.0f kcyc" % tuple(dbg[:, i].mean() / 1e3 for i in (0, 1, 2, 4, 5, 6)))
                print("        warp4 store %6.0f fence %6.0f | warp8 fence %6.0f sync %6.0f kcyc" % tuple(
                    (g[:, 12 + i].mean() / 1e3) for i in range(4)))
        del b
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
