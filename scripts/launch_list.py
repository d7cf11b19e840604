#!/usr/bin/env python3
"""Launches per kernel tag of one fwd+bwd of every block at a given batch (automatic path selection):
   python scripts/launch_list.py --n 32"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import SHAPES, BlockBuffers  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
a = ap.parse_args()
lib = L.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream()
total = 0
for c, h in SHAPES:
    b = BlockBuffers(torch, L, a.n, c, h, dev, seed=c)
    b.fwd_bwd(lib, L, st.cuda_stream, 0)
    torch.cuda.synchronize()
    lib.gml_profile_reset()
    lib.gml_profile_enable(1)
    b.fwd_bwd(lib, L, st.cuda_stream, 0)
    torch.cuda.synchronize()
    lib.gml_profile_enable(0)
    row = []
    for tag in range(lib.gml_kernel_tag_count()):
        tot, cnt = ctypes.c_double(), ctypes.c_int64()
        lib.gml_profile_read(tag, ctypes.byref(tot), ctypes.byref(cnt))
        if cnt.value:
            row.append("%s x%d (%.1f us)" % (lib.gml_kernel_tag_name(tag).decode(), cnt.value, tot.value * 1e3))
            total += cnt.value
    print("%dx%d^2 n=%d: %s" % (c, h, a.n, ", ".join(row)))
print("launches per step:", total)
