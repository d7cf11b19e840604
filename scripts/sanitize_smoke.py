#!/usr/bin/env python3
"""Tiny driver for compute-sanitizer: every kernel path once on small shapes, checked against the oracle.
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import greedy_multimodal_learning_b200 as pkg  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402
from oracle import mmtm_oracle as mo  # noqa: E402

dev = "cuda:0"
lib = L.load()


def run(n, c, h, w, mode, flags, tun):
    for k, v in {"fused_kind": 0, "fused_cluster": 0, "fused_threads": 0, **tun}.items():
        L.check(lib.gml_set_tunable(k.encode(), v))
    x = mo.synth_inputs(n + c, n, c, h, w)
    p = mo.synth_params(c, c, c)
    m = pkg.MMTM_mitigate(c, c, 4, kernel_flags=flags)
    with torch.no_grad():
        for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                             m.fc_skeleton.weight, m.fc_skeleton.bias), p.tensors()):
            dst.copy_(src)
    m.to(dev)
    avg = [torch.zeros(c) + 0.1, torch.zeros(c) - 0.1]
    kw = {1: dict(curation_mode=True, caring_modality=0), 2: dict(curation_mode=True, caring_modality=1),
          3: dict(turnoff_cross_modal_flow=True, average_squeezemaps=[a.to(dev) for a in avg])}.get(mode, {})
    a = x["A"].to(dev).requires_grad_(True)
    b = x["B"].to(dev).requires_grad_(True)
    a_out, b_out, _, _ = m(a, b, **kw)
    torch.autograd.backward([a_out, b_out], [x["gA"].to(dev), x["gB"].to(dev)])
    o = mo.forward_backward(x["A"], x["B"], p, mo.MMTMState.zeros(c), x["gA"], x["gB"], mode, avg)
    err = max(float((a_out.detach().cpu() - o["A_out"]).abs().max()), float((a.grad.cpu() - o["dA"]).abs().max()))
    assert err < 1e-4, err
    return err


cases = []
for mode in range(4):
    cases.append((3, 12, 3, 5, mode, L.F_FORCE_STREAMING, {}))           # scalar path, odd planes
    cases.append((5, 16, 4, 4, mode, L.F_FORCE_STREAMING, {}))           # vector path
for kind, cs, thr in ((1, 4, 512), (1, 4, 256), (1, 8, 256), (2, 8, 0), (2, 4, 0)):
    for shape in ((3, 32, 8, 8), (5, 64, 6, 6), (9, 128, 28, 28)):
        cases.append(shape + (0, L.F_FORCE_FUSED, {"fused_kind": kind, "fused_cluster": cs, "fused_threads": thr}))
worst = 0.0
for cse in cases:
    try:
        worst = max(worst, run(*cse))
    except L.GmlError as e:
        if "unsupported" not in str(e):
            raise
# statistics kernels
lin = torch.nn.Linear(301, 77).to(dev)
lin(torch.randn(4, 301, device=dev)).sum().backward()
sq = pkg.MultiTensorSqnorm([("net_view_0.w", lin.weight), ("mmtm.fc_visual.b", lin.bias)], ["net_view_0", "net_view_1"],
                           ["visual", "skeleton"])
sq.measure()
counts = torch.empty(3, dtype=torch.int32, device=dev)
l0, l1 = torch.randn(9, 40, device=dev), torch.randn(9, 40, device=dev)
y = torch.randint(0, 40, (9,), device=dev)
L.check(lib.gml_accuracy_counts(l0.data_ptr(), l1.data_ptr(), y.data_ptr(), 9, 40, counts.data_ptr(),
                                L.current_stream(torch.device(dev))))
s = torch.zeros(40, dtype=torch.float64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
L.check(lib.gml_squeeze_accumulate(l0.data_ptr(), None, 9, 40, s.data_ptr(), cnt.data_ptr(),
                                   L.current_stream(torch.device(dev))))
torch.cuda.synchronize()
print("sanitize smoke ok: %d MMTM cases, worst abs error %.2e" % (len(cases), worst))
