#!/usr/bin/env python3
"""One-launch learning-speed reduction over the real model (ncu target):  python scripts/stats_one.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import greedy_multimodal_learning_b200 as pkg  # noqa: E402

torch.manual_seed(777)
dev = torch.device("cuda:0")
model = pkg.MMTM_MVCNN().to(dev)
for p in model.parameters():
    p.grad = torch.randn_like(p) * 0.01
sq = pkg.MultiTensorSqnorm(model.named_parameters(), ["net_view_0", "net_view_1"], ["visual", "skeleton"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    r = sq.measure()
print("ok", r["gn_main"])
