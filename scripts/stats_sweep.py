#!/usr/bin/env python3
"""Device time of the one-launch learning-speed reduction over the real model for each scheduling variant
(tunable sq_variant), cold L2:  python scripts/stats_sweep.py"""
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import greedy_multimodal_learning_b200 as pkg  # noqa: E402
from greedy_multimodal_learning_b200 import _lib  # noqa: E402

torch.manual_seed(777)
dev = torch.device("cuda:0")
lib = _lib.load()
model = pkg.MMTM_MVCNN().to(dev)
tensors = []
for p in model.parameters():
    p.grad = torch.randn_like(p) * 0.01
    tensors += [p.detach(), p.grad]
n = len(tensors)
ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
numel = (ctypes.c_int64 * n)(*[t.numel() for t in tensors])
m = (ctypes.c_int32 * n)(*([1] * n))
k = (ctypes.c_int32 * n)(*[i % 2 for i in range(n)])
ws_bytes = lib.gml_sqnorm_workspace_bytes(numel, n)
ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
out = torch.zeros(8, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nbytes = sum(t.numel() * 4 for t in tensors)
st = _lib.current_stream(dev)
ref = None
names = {0: "4096", 1: "2048", 2: "1024", 3: "8192"}
clean = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for drain, bps, v in [(d, b, v) for d in (False, True) for b in (0, 6) for v in (0, 1, 2, 3, 4)]:
    if True:
        var = v | (bps << 3)
        _lib.check(lib.gml_set_tunable(b"sq_variant", var))
        ts = []
        for i in range(25):
            flush.zero_()
            if drain:
                # the 256 MB just written leave ~100 MB of DIRTY lines in L2; their write-back competes with the
                # scan for HBM.  Reading another 256 MB drains them and leaves L2 full of clean, unrelated lines.
                clean.view(torch.int64).sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.gml_multi_tensor_sqnorm(ptrs, numel, m, k, n, out.data_ptr(), None, ws.data_ptr(), ws_bytes, st))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        o = out.cpu().numpy().copy()
        if ref is None:
            ref = o
        rel = float(abs(o - ref).max() / abs(ref).max())
        ts = ts[5:]
        med = statistics.median(ts)
        print("%s chunk %s %-7s blocks/SM %d: median %.1f us min %.1f us  -> %.0f GB/s (%.3f of 6550)  rel diff vs first %.1e"
              % ("drained" if drain else "dirty  ", names[v & 3], "no-pdl" if v & 4 else "pdl", bps or 8, med, min(ts), nbytes / med / 1e3,
                 nbytes / med / 1e3 / 6550.1, rel))
_lib.check(lib.gml_set_tunable(b"sq_variant", 0))
