#!/usr/bin/env python3
"""Errors of every output of the (1100, 512, 1x1) block vs the oracle, per GEMM variant and mode."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from greedy_multimodal_learning_b200 import _lib  # noqa: E402
from oracle import mmtm_oracle as mo  # noqa: E402
from tests.helpers import rel_err  # noqa: E402
from tests.test_mmtm_gpu import PATHS, make_module, run_cuda  # noqa: E402

n, c_v, c_s, h_v, w_v, h_s, w_s = (1100, 512, 512, 1, 1, 1, 1)
lib = _lib.load()
for mode in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,3").split(",")]:
    for umma in (0, 1):
        _lib.check(lib.gml_set_tunable(b"gemm_umma", umma))
        rs = np.random.RandomState(n * 1000 + c_v)
        t = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
        x = dict(A=t(n, c_v, h_v, w_v), B=t(n, c_s, h_s, w_s), gA=t(n, c_v, h_v, w_v), gB=t(n, c_s, h_s, w_s))
        warm = dict(A=t(2, c_v, h_v, w_v), B=t(2, c_s, h_s, w_s))
        p = mo.synth_params(5, c_v, c_s)
        avg = [0.1 * t(c_v), 0.1 * t(c_s)]
        m = make_module(c_v, c_s, p, PATHS["streaming"])
        r = run_cuda(m, x, mode, avg, warm)
        st = mo.MMTMState.zeros(c_v)
        with torch.no_grad():
            mo.forward(warm["A"], warm["B"], p, st, 0)
        o = mo.forward_backward(x["A"], x["B"], p, st, x["gA"], x["gB"], mode, avg)
        print("mode %d umma %d: %s" % (mode, umma, "  ".join(
            "%s %.1e" % (k, rel_err(r[k], o[k])) for k in ["A_out", "B_out", "gA", "gB", "dA", "dB", "dWsq", "dbsq", "dWv",
                                                          "dbv", "dWs", "dbs"])), flush=True)
