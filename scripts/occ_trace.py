#!/usr/bin/env python3
"""Residency timeline of the L2-resident cluster kernels: every CTA records {smid, start ns, end ns} (globaltimer);
this script reports the kernel span, CTA life, average resident CTAs per SM and the idle time between consecutive
CTAs of an SM slot.  Measurement only (tunable fused_occ_trace_ptr)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import BlockBuffers  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
N = int(os.environ.get("TRACE_N", "1024"))
c, h = 128, 28
b = BlockBuffers(torch, L, N, c, h, dev, seed=c)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P = lambda t: t.data_ptr()
w, dw = b.w, b.dw


def fwd():
    L.check(lib.gml_mmtm_fwd(P(b.a), P(b.b), P(b.a_out), P(b.b_out), P(w[0]), P(w[1]), P(w[2]), P(w[3]), P(w[4]), P(w[5]),
                             P(b.z), P(b.hid), P(b.g_a), P(b.g_b), P(b.gate_sum), P(b.run_v), P(b.run_s), 0, None, None,
                             P(b.fws), b.fws_bytes, b.dims, 0, 1.0, 0, st))


def bwd():
    L.check(lib.gml_mmtm_bwd(P(b.go_a), P(b.go_b), P(b.a), P(b.b), P(w[0]), P(w[2]), P(w[4]), P(b.z), P(b.hid), P(b.g_a),
                             P(b.g_b), None, None, None, None, P(b.d_a), P(b.d_b), P(dw[0]), P(dw[1]), P(dw[2]), P(dw[3]),
                             P(dw[4]), P(dw[5]), P(b.ws), b.ws_bytes, b.dims, 0, 1.0, 0, st))


for what, fn in (("fwd", fwd), ("bwd", bwd)):
    fn()
    torch.cuda.synchronize()
    n_cta = 8 * N * 2
    trace = torch.zeros(n_cta * 4, dtype=torch.int64, device=dev)
    flush.zero_()
    torch.cuda.synchronize()
    L.check(lib.gml_set_tunable(b"fused_occ_trace_ptr", trace.data_ptr()))
    fn()
    torch.cuda.synchronize()
    L.check(lib.gml_set_tunable(b"fused_occ_trace_ptr", 0))
    t = trace.cpu().numpy().reshape(-1, 4)
    t = t[t[:, 1] > 0]
    sm, t0, t1 = t[:, 0], t[:, 1].astype(np.float64), t[:, 2].astype(np.float64)
    base = t0.min()
    t0 -= base
    t1 -= base
    span = t1.max()
    life = t1 - t0
    print("== C=%d H=%d N=%d %s: %d CTAs on %d SMs, span %.1f us" % (c, h, N, what, len(t), len(set(sm.tolist())), span / 1e3))
    print("   CTA life us: mean %.2f  p10 %.2f  p50 %.2f  p90 %.2f  max %.2f" %
          (life.mean() / 1e3, np.percentile(life, 10) / 1e3, np.percentile(life, 50) / 1e3, np.percentile(life, 90) / 1e3,
           life.max() / 1e3))
    print("   average resident CTAs per SM over the span: %.2f (sum of lives / span / SMs)" %
          (life.sum() / span / len(set(sm.tolist()))))
    # residency histogram: sample the timeline
    grid = np.linspace(0, span, 400)
    res = np.zeros_like(grid)
    for i, g in enumerate(grid):
        res[i] = np.count_nonzero((t0 <= g) & (t1 > g))
    nsm = len(set(sm.tolist()))
    print("   resident CTAs per SM along the span (10 deciles): " + " ".join("%.2f" % (res[k * 40:(k + 1) * 40].mean() / nsm) for k in range(10)))
    # start-time clustering: how long after a CTA of an SM ends does the next one start on that SM
    gaps = []
    for s_ in set(sm.tolist()):
        m = sm == s_
        ends = np.sort(t1[m])
        starts = np.sort(t0[m])
        # match every start after the first wave with the latest end before it
        for x in starts[4:]:
            k = np.searchsorted(ends, x, side="right") - 1
            if k >= 0:
                gaps.append(x - ends[k])
    gaps = np.array(gaps)
    print("   start - latest earlier end on the same SM (us): mean %.2f p50 %.2f p90 %.2f" %
          (gaps.mean() / 1e3, np.percentile(gaps, 50) / 1e3, np.percentile(gaps, 90) / 1e3))
    # first-wave start skew
    print("   first 592 CTA starts (us): p50 %.2f p90 %.2f max %.2f" %
          (np.percentile(np.sort(t0)[:592], 50) / 1e3, np.percentile(np.sort(t0)[:592], 90) / 1e3, np.sort(t0)[:592].max() / 1e3))
