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
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--shapes", default="")
    args = ap.parse_args()
    lib = L.load()
    dev = torch.device("cuda:0")
    for kv in filter(None, args.tunables.split(",")):
        k, v = kv.split("=")
        L.check(lib.gml_set_tunable(k.encode(), int(v)))
    stats = torch.zeros(1024 * 16, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for c, h in SHAPES:
        if args.shapes and str(c) not in args.shapes.split(","):
            continue
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
            big = torch.iinfo(torch.int64).max
            stats.view(-1, 16)[399, 0] = big
            stats.view(-1, 16)[700:764, 0] = big
            stats.view(-1, 16)[700:764, 3] = big
            L.check(lib.gml_set_tunable(b"tile_trace_only", 1 if args.timeline else 0))
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
            if len(r):
                print("        loader per item (cyc): slot-wait %5.0f dep-wait %5.0f head(segment/loop) %5.0f issue(meta+TMA) %5.0f ticket-draw wait %5.0f" % tuple(
                    float(r[:, i].sum() / r[:, 1].sum()) for i in (2, 3, 8, 9, 10)))
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
            if args.timeline:
                t0 = float(s_all[399, 0])
                us = lambda v: (float(v) - t0) / 1e3
                print("        tile timeline (us from kernel start): R first issue | R last publish | S gates seen (last CTA) | S first issue | S last issue")
                for t in range(64):
                    row = s_all[700 + t]
                    if row[1] == 0:
                        continue
                    print("          tile %2d: %7.1f %7.1f %7.1f %7.1f %7.1f" % (t, us(row[0]), us(row[1]), us(row[2]), us(row[3]), us(row[4])))
                print("        GEMM items: ticket tile stage ntile cta | drawn  deps-ok  main-done  signalled (us)")
                for k in range(256):
                    row = s_all[400 + k]
                    if row[5] == 0:
                        continue
                    print("          %3d  t%-2d s%d n%-2d cta%-3d | %7.1f %7.1f %7.1f %7.1f" % (
                        k, row[0], row[1], row[2], row[3], us(row[4]), us(row[5]), us(row[6]), us(row[7])))
        del b
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
