#!/usr/bin/env python3
"""tcgen05 GEMM on a few shapes and both operand layouts: errors vs float64 and, with --trace, the per-phase
globaltimer stamps of CTA (0,0,0) (slots: 0 start, 1 TMEM ready, 2+4i tile i landed, 3+4i tile i split and
published, 4+4i drain done, 40 loads drained, 41 accumulators drained, 42 partial tile parked, 43 cluster
barrier, 44 reduced and stored, 45 end)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402
from tests.test_gemm_gpu import _operand  # noqa: E402

lib, dev = L.load(), torch.device("cuda:0")
ws_bytes = lib.gml_fc_gemm_workspace_bytes()
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
L.check(lib.gml_set_tunable(b"gemm_umma", 1))
if True:
    for (m, n, k, a_kc, b_kc) in [(256, 256, 128, 1, 1), (256, 256, 128, 1, 0), (256, 256, 128, 0, 1), (256, 256, 128, 0, 0),
                                  (1024, 512, 1024, 1, 1), (512, 1024, 1024, 0, 0)]:
        rs = np.random.RandomState(1)
        a, lda, a64 = _operand(rs, m, k, a_kc, dev)
        b, ldb, b64 = _operand(rs, n, k, b_kc, dev)
        c = torch.zeros(m, n, device=dev)
        L.check(lib.gml_fc_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, m, n, k, lda, ldb, n, a_kc, b_kc, 0, 0,
                                ws.data_ptr(), ws_bytes, st))
        got = c.cpu().numpy().astype(np.float64)
        want = a64 @ b64.T
        if "--trace" in sys.argv:
            tr = torch.zeros(64, dtype=torch.int64, device=dev)
            L.check(lib.gml_set_tunable(b"gemm_trace_ptr", tr.data_ptr()))
            for _ in range(3):
                lib.gml_fc_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, m, n, k, lda, ldb, n, a_kc, b_kc, 0, 0,
                                ws.data_ptr(), ws_bytes, st)
            torch.cuda.synchronize()
            L.check(lib.gml_set_tunable(b"gemm_trace_ptr", 0))
            t = tr.cpu().numpy()
            print("   trace (ns from start):", " ".join("%d:%d" % (i, t[i] - t[0]) for i in range(64) if t[i]))
        print("m%d n%d k%d a_kc=%d b_kc=%d  max err %.3e  (|want| max %.1f)" % (
            m, n, k, a_kc, b_kc, np.abs(got - want).max(), np.abs(want).max()), flush=True)
