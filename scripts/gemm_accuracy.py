#!/usr/bin/env python3
"""Accuracy and speed of the FC GEMM variants against float64 (and torch fp32 matmul as the yardstick)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402
from tests.test_gemm_gpu import SHAPES, VARIANTS, _operand  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
lib, dev = L.load(), torch.device("cuda:0")
ws_bytes = lib.gml_fc_gemm_workspace_bytes()
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
print("%-28s %10s | %s" % ("shape", "torch fp32", " | ".join("%-22s" % v for v in VARIANTS)))
for shape in SHAPES:
    m, n, k, a_kc, b_kc = shape
    rs = np.random.RandomState(m * 7 + n * 3 + k)
    a, lda, a64 = _operand(rs, m, k, a_kc, dev)
    b, ldb, b64 = _operand(rs, n, k, b_kc, dev)
    want = a64 @ b64.T
    ta = torch.from_numpy(a64.astype(np.float32)).to(dev)
    tb = torch.from_numpy(b64.astype(np.float32)).to(dev)
    e_torch = np.abs((ta @ tb.T).cpu().numpy().astype(np.float64) - want).max()
    cells = []
    for name, tun in VARIANTS.items():
        for key, val in tun.items():
            L.check(lib.gml_set_tunable(key.encode(), val))
        c = torch.zeros(m, n, device=dev)
        call = lambda: lib.gml_fc_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, m, n, k, lda, ldb, n, a_kc, b_kc,
                                       0, 0, ws.data_ptr(), ws_bytes, st)
        L.check(call())
        err = np.abs(c.cpu().numpy().astype(np.float64) - want).max()
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        cells.append("%.2e %6.1fus %5.1fTF" % (err, us, 2.0 * m * n * k / us / 1e6))
    print("m%dn%dk%d_%d%d %s %10.2e | %s" % (m, n, k, a_kc, b_kc, " " * (28 - len("m%dn%dk%d_%d%d" % shape) - 1), e_torch,
                                             " | ".join(cells)), flush=True)
