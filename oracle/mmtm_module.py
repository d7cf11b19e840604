"""nn.Module wrapper around the oracle MMTM (oracle/mmtm_oracle.py).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Gives the oracle the same constructor / forward
surface as the reference's `MMTM_mitigate` so that
  * CPU tests can plug it into the host-side mirrors (model, step engine, callbacks) and
    replay the reference's recorded training trace without a GPU, and
  * `bench.py` can time a full CPU reference-path train step as the `cpu_baseline`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import mmtm_oracle as mo


class OracleMMTM(nn.Module):
    def __init__(self, dim_visual, dim_skeleton, ratio, device=0, SEonly=False, shareweight=False):
        super().__init__()
        assert not SEonly and not shareweight
        d = mo.hidden_dim(dim_visual, dim_skeleton, ratio)
        self.dim_visual, self.dim_skeleton = dim_visual, dim_skeleton
        self.state = mo.MMTMState.zeros(dim_visual)
        self.fc_squeeze = nn.Linear(dim_visual + dim_skeleton, d)
        self.fc_visual = nn.Linear(d, dim_visual)
        self.fc_skeleton = nn.Linear(d, dim_skeleton)

    @property
    def running_avg_weight_visual(self):
        return self.state.run_v

    @property
    def running_avg_weight_skeleton(self):
        return self.state.run_s

    @property
    def step(self):
        return self.state.step

    def forward(self, visual, skeleton, return_scale=False, return_squeezed_mps=False,
                turnoff_cross_modal_flow=False, average_squeezemaps=None, curation_mode=False, caring_modality=0):
        mode = mo.mode_from_flags(curation_mode, caring_modality, turnoff_cross_modal_flow)
        p = mo.MMTMParams.from_module(self)
        if self.state.run_v.device != visual.device:
            self.state.run_v = self.state.run_v.to(visual.device)
            self.state.run_s = self.state.run_s.to(visual.device)
        a_out, b_out, aux = mo.forward(visual, skeleton, p, self.state, mode, average_squeezemaps)
        scales = [aux["gA"].detach().cpu(), aux["gB"].detach().cpu()] if return_scale else None
        sq = [aux["sA"].detach().cpu(), aux["sB"].detach().cpu()] if return_squeezed_mps else None
        return a_out, b_out, scales, sq
