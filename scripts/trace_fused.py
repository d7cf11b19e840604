#!/usr/bin/env python3
"""Per-phase clock64() timeline of the fused kernels (first 8 CTAs, first 16 groups each)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import BlockBuffers  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
NAMES = ["pass1", "sync", "scatter+csync", "fc1", "csync", "fc2", "sync", "pass2", "sync"]
KIND = int(os.environ.get("TRACE_KIND", "1"))
SHAPES = [(128, 28), (256, 14)] if not os.environ.get("TRACE_C") else [(int(os.environ["TRACE_C"]), {128: 28, 256: 14}[int(os.environ["TRACE_C"])])]
for (c, h, n) in [(c_, h_, int(os.environ.get("TRACE_N", "256"))) for c_, h_ in SHAPES]:
    for cs, thr in (((4, 512), (8, 256)) if KIND == 1 else ((8, 0), (4, 0))):
        L.check(lib.gml_set_tunable(b"fused_kind", KIND))
        L.check(lib.gml_set_tunable(b"fused_cluster", cs))
        L.check(lib.gml_set_tunable(b"fused_threads", thr))
        b = BlockBuffers(torch, L, n, c, h, dev, seed=c)
        st = torch.cuda.current_stream().cuda_stream
        for what in ("fwd", "bwd"):
            trace = torch.zeros(8 * 16 * 16, dtype=torch.int64, device=dev)
            try:
                b.fwd_bwd(lib, L, st, L.F_FORCE_FUSED)  # warm
            except L.GmlError:
                print("== C=%d cs=%d unsupported" % (c, cs))
                break
            torch.cuda.synchronize()
            L.check(lib.gml_set_tunable(b"fused_trace_ptr", trace.data_ptr()))
            if what == "fwd":
                # only forward traced: run fwd+bwd but the bwd overwrites -> trace fwd via a fwd-only call
                P = lambda t: t.data_ptr()
                w = b.w
                L.check(lib.gml_mmtm_fwd(P(b.a), P(b.b), P(b.a_out), P(b.b_out), P(w[0]), P(w[1]), P(w[2]), P(w[3]),
                                         P(w[4]), P(w[5]), P(b.z), P(b.hid), P(b.g_a), P(b.g_b), P(b.gate_sum),
                                         P(b.run_v), P(b.run_s), 0, None, None, P(b.fws), b.fws_bytes, b.dims, 0, 1.0,
                                         L.F_FORCE_FUSED, st))
            else:
                L.check(lib.gml_set_tunable(b"fused_trace_ptr", 0))
                P = lambda t: t.data_ptr()
                w, dw = b.w, b.dw
                L.check(lib.gml_set_tunable(b"fused_trace_ptr", trace.data_ptr()))
                L.check(lib.gml_mmtm_bwd(P(b.go_a), P(b.go_b), P(b.a), P(b.b), P(w[0]), P(w[2]), P(w[4]), P(b.z),
                                         P(b.hid), P(b.g_a), P(b.g_b), None, None, None, None, P(b.d_a), P(b.d_b),
                                         P(dw[0]), P(dw[1]), P(dw[2]), P(dw[3]), P(dw[4]), P(dw[5]), P(b.ws),
                                         b.ws_bytes, b.dims, 0, 1.0, L.F_FORCE_FUSED, st))
            torch.cuda.synchronize()
            L.check(lib.gml_set_tunable(b"fused_trace_ptr", 0))
            t = trace.cpu().view(8, 16, 16).numpy()
            print("== C=%d H=%d N=%d cs=%d T=%d %s (cycles; CTA 0 and CTA 5)" % (c, h, n, cs, thr, what))
            for cta in ((0, 5) if KIND == 1 else (0, 1, 2, 5)):
                for it in range(8 if KIND == 1 else 1):
                    row = t[cta, it]
                    if row[0] == 0:
                        break
                    d = [int(row[k + 1] - row[k]) for k in range(9)]
                    gap = int(row[0] - t[cta, it - 1][9]) if it else 0
                    print("  cta%d it%d total %6d | " % (cta, it, int(row[9] - row[0])) +
                          " ".join("%s=%d" % (nm, v) for nm, v in zip(NAMES, d)) + " | gap %d" % gap)
        del b
L.check(lib.gml_set_tunable(b"fused_cluster", 0))
L.check(lib.gml_set_tunable(b"fused_threads", 0))
