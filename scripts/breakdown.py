#!/usr/bin/env python3
"""Per-kernel-class CUDA-event breakdown of one MMTM block's forward and backward (C ABI)."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import SHAPES, BlockBuffers  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402
from scripts.sweep import VARIANTS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="256")
ap.add_argument("--variants", default="stream_nochunk,fused_cs8_t256")
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
lib = L.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for c, h in SHAPES:
    for n in [int(x) for x in args.batches.split(",")]:
        b = BlockBuffers(torch, L, n, c, h, dev, seed=c)
        P = lambda t: t.data_ptr()
        w, dw = b.w, b.dw
        fwd = lambda fl: lib.gml_mmtm_fwd(P(b.a), P(b.b), P(b.a_out), P(b.b_out), P(w[0]), P(w[1]), P(w[2]), P(w[3]),
                                          P(w[4]), P(w[5]), P(b.z), P(b.hid), P(b.g_a), P(b.g_b), P(b.gate_sum),
                                          P(b.run_v), P(b.run_s), 0, None, None, P(b.fws), b.fws_bytes, b.dims, 0, 1.0,
                                          fl, st)
        bwd = lambda fl: lib.gml_mmtm_bwd(P(b.go_a), P(b.go_b), P(b.a), P(b.b), P(w[0]), P(w[2]), P(w[4]), P(b.z),
                                          P(b.hid), P(b.g_a), P(b.g_b), None, None, None, None, P(b.d_a), P(b.d_b),
                                          P(dw[0]), P(dw[1]), P(dw[2]), P(dw[3]), P(dw[4]), P(dw[5]), P(b.ws),
                                          b.ws_bytes, b.dims, 0, 1.0, fl, st)
        for name in args.variants.split(","):
            flags, tun = VARIANTS[name]
            for k, v in {"l2_chunk_mb": 100000, "fused_kind": 0, "fused_cluster": 0, "fused_threads": 0, "gemm_big_tiles": 0, "gemm_tf32x3": 1, "gemm_umma": 1, **tun}.items():
                L.check(lib.gml_set_tunable(k.encode(), v))
            if fwd(flags) == -5:
                continue
            L.check(bwd(flags))
            torch.cuda.synchronize()
            for what, fn in (("fwd", fwd), ("bwd", bwd)):
                lib.gml_profile_reset()
                for _ in range(args.iters):
                    flush.zero_()
                    torch.cuda.synchronize()
                    lib.gml_profile_enable(1)
                    L.check(fn(flags))
                    lib.gml_profile_enable(0)
                parts = []
                for tag in range(lib.gml_kernel_tag_count()):
                    tot, cnt = ctypes.c_double(), ctypes.c_int64()
                    lib.gml_profile_read(tag, ctypes.byref(tot), ctypes.byref(cnt))
                    if cnt.value:
                        parts.append("%s %.1fus x%d" % (lib.gml_kernel_tag_name(tag).decode(),
                                                       1e3 * tot.value / args.iters, cnt.value // args.iters))
                print("C=%3d H=%2d N=%4d %-15s %s: %s" % (c, h, n, name, what, " | ".join(parts)), flush=True)
        del b
        torch.cuda.empty_cache()
