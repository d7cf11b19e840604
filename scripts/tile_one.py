#!/usr/bin/env python3
"""Run the forced tile-pipeline forward (and optionally backward) of one block a few times (ncu target).
   python scripts/tile_one.py --c 128 --h 28 --n 256 [--bwd] [--reps 3] [--tunables k=v,...]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import BlockBuffers  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=128)
ap.add_argument("--h", type=int, default=28)
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--path", default="tile")
ap.add_argument("--tunables", default="")
a = ap.parse_args()
lib = L.load()
for kv in filter(None, a.tunables.split(",")):
    k, v = kv.split("=")
    L.check(lib.gml_set_tunable(k.encode(), int(v)))
dev = torch.device("cuda:0")
b = BlockBuffers(torch, L, a.n, a.c, a.h, dev, seed=1)
flags = {"tile": L.F_FORCE_TILE, "old": 0, "auto": 0}[a.path]
if a.path == "old":
    L.check(lib.gml_set_tunable(b"tile_kind", 2))
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
w, dw = b.w, b.dw
for _ in range(a.reps):
    L.check(lib.gml_mmtm_fwd(P(b.a), P(b.b), P(b.a_out), P(b.b_out), P(w[0]), P(w[1]), P(w[2]), P(w[3]), P(w[4]), P(w[5]),
                             P(b.z), P(b.hid), P(b.g_a), P(b.g_b), P(b.gate_sum), P(b.run_v), P(b.run_s), 0, None, None,
                             P(b.fws), b.fws_bytes, b.dims, 0, 1.0, flags, st))
    if a.bwd:
        L.check(lib.gml_mmtm_bwd(P(b.go_a), P(b.go_b), P(b.a), P(b.b), P(w[0]), P(w[2]), P(w[4]), P(b.z), P(b.hid),
                                 P(b.g_a), P(b.g_b), None, None, None, None, P(b.d_a), P(b.d_b), P(dw[0]), P(dw[1]),
                                 P(dw[2]), P(dw[3]), P(dw[4]), P(dw[5]), P(b.ws), b.ws_bytes, b.dims, 0, 1.0, flags, st))
torch.cuda.synchronize()
print("ok", float(b.a_out.abs().sum()))
