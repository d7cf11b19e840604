"""CPU oracle for the MMTM fusion block (forward, backward, running gate mean, modes).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this.
The product path (`greedy_multimodal_learning_b200`) never routes through it.

This is a from-scratch restatement of the algorithm in the reference's
`src/balanced_mmtm.py:93-154` (class `MMTM_mitigate.forward`), written functionally
over explicit weight matrices.  PARITY STATUS: **pinned** -- the reference has no
tests or golden vectors of its own, so the restatement is
pinned against outputs of the *unmodified reference executed in the authoring
container*: `tests/golden/make_golden.py` (committed) imports `/root/reference`
and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this file
against those vectors, and `tests/test_oracle_vs_reference.py` re-runs the live
reference whenever `/root/reference` is present.

Two implementations live here:
  * `forward()` / `forward_backward()`  -- torch CPU fp32, autograd backward: the same
    arithmetic class as the reference's eager path; this is the "port" that
    `bench.py` times as the CPU baseline.
  * `forward_backward_f64()`            -- numpy float64, closed-form backward derived by
    hand: an independent check of the gradient formulas the CUDA kernels implement.

Mode numbering (shared with include/gml_b200.h):
  0 normal                         balanced_mmtm.py:93-111,128-133
  1 curation, caring_modality==0   balanced_mmtm.py:135-143  (visual gate <- running mean)
  2 curation, caring_modality==1   balanced_mmtm.py:145-152  (skeleton gate <- running mean)
  3 cross-modal flow off           balanced_mmtm.py:72-91
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

MODE_NORMAL = 0
MODE_CURATE_VISUAL = 1    # caring_modality == 0: visual gate replaced by its running mean
MODE_CURATE_SKELETON = 2  # caring_modality == 1: skeleton gate replaced by its running mean
MODE_XMODAL_OFF = 3


def mode_from_flags(curation_mode: bool, caring_modality, turnoff_cross_modal_flow: bool) -> int:
    """Map the reference's three forward kwargs onto one mode id.

    balanced_mmtm.py:128-152 -- `curation_mode` False => plain gating.  With
    curation on, caring_modality 0 / 1 pick the substituted side; any other value
    (e.g. None) leaves BOTH gates as un-reshaped [N, C] tensors, which then fail to
    broadcast against [N, C, H, W] -- the reference raises, so do we.
    Cross-modal-off only changes how the gates are computed (:72-91); curation can in
    principle be layered on top, the reference never does (eval.gin) so we keep
    mode 3 exclusive of curation.
    """
    if turnoff_cross_modal_flow:
        if curation_mode:
            raise NotImplementedError("curation on top of cross-modal-off is never used by the reference")
        return MODE_XMODAL_OFF
    if not curation_mode:
        return MODE_NORMAL
    if caring_modality == 0:
        return MODE_CURATE_VISUAL
    if caring_modality == 1:
        return MODE_CURATE_SKELETON
    raise ValueError("curation_mode=True needs caring_modality in {0, 1}")


@dataclass
class MMTMParams:
    """Weights of one block, reference names in brackets (balanced_mmtm.py:37-45)."""
    w_sq: torch.Tensor  # [D, Cv+Cs]  fc_squeeze.weight
    b_sq: torch.Tensor  # [D]         fc_squeeze.bias
    w_v: torch.Tensor   # [Cv, D]     fc_visual.weight
    b_v: torch.Tensor   # [Cv]        fc_visual.bias
    w_s: torch.Tensor   # [Cs, D]     fc_skeleton.weight
    b_s: torch.Tensor   # [Cs]        fc_skeleton.bias

    def tensors(self):
        return (self.w_sq, self.b_sq, self.w_v, self.b_v, self.w_s, self.b_s)

    @staticmethod
    def from_module(m) -> "MMTMParams":
        return MMTMParams(m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight,
                          m.fc_visual.bias, m.fc_skeleton.weight, m.fc_skeleton.bias)

    def detach_clone(self, requires_grad=False) -> "MMTMParams":
        return MMTMParams(*[t.detach().clone().requires_grad_(requires_grad) for t in self.tensors()])


@dataclass
class MMTMState:
    """Non-persistent running statistics (balanced_mmtm.py:30-32)."""
    run_v: torch.Tensor
    run_s: torch.Tensor
    step: int = 0

    @staticmethod
    def zeros(dim_visual: int) -> "MMTMState":
        # NB both buffers are sized by dim_visual in the reference (:30-31).
        return MMTMState(torch.zeros(dim_visual), torch.zeros(dim_visual), 0)

    def clone(self) -> "MMTMState":
        return MMTMState(self.run_v.clone(), self.run_s.clone(), self.step)


def hidden_dim(dim_visual: int, dim_skeleton: int, ratio) -> int:
    """balanced_mmtm.py:25-26."""
    return int(2 * (dim_visual + dim_skeleton) / ratio)


def _plane_mean(x: torch.Tensor) -> torch.Tensor:
    # balanced_mmtm.py:96-97: view(N, C, -1) then mean over the last axis.
    return x.reshape(x.shape[0], x.shape[1], -1).mean(dim=-1)


def gates(a: torch.Tensor, b: torch.Tensor, p: MMTMParams, mode: int,
          avg: Optional[Sequence[torch.Tensor]] = None):
    """Squeeze + excitation.  Returns (sA, sB, gA, gB, hidden list)."""
    s_a, s_b = _plane_mean(a), _plane_mean(b)
    n = a.shape[0]
    lin = torch.nn.functional.linear
    if mode == MODE_XMODAL_OFF:
        # balanced_mmtm.py:72-91: each modality sees the dataset-mean squeeze of the other.
        m_a, m_b = avg[0], avg[1]
        z1 = torch.cat([s_a, m_b.unsqueeze(0).expand(n, -1)], dim=1)
        z2 = torch.cat([m_a.unsqueeze(0).expand(n, -1), s_b], dim=1)
        h1 = torch.relu(lin(z1, p.w_sq, p.b_sq))
        h2 = torch.relu(lin(z2, p.w_sq, p.b_sq))
        e_a = lin(h1, p.w_v, p.b_v)
        e_b = lin(h2, p.w_s, p.b_s)
        hs = [h1, h2]
    else:
        # balanced_mmtm.py:93-109: visual first in the concat.
        z = torch.cat([s_a, s_b], dim=1)
        h = torch.relu(lin(z, p.w_sq, p.b_sq))
        e_a = lin(h, p.w_v, p.b_v)
        e_b = lin(h, p.w_s, p.b_s)
        hs = [h]
    # balanced_mmtm.py:110-111 -- plain sigmoid (NOT 2*sigmoid).
    return s_a, s_b, torch.sigmoid(e_a), torch.sigmoid(e_b), hs


def update_running(state: MMTMState, g_a: torch.Tensor) -> None:
    """balanced_mmtm.py:113-116.  BOTH running means are fed by the visual gate (sic)."""
    mean_ga = g_a.detach().mean(dim=0)
    k = state.step
    state.run_v = (mean_ga + state.run_v * k) / (k + 1)
    state.run_s = (mean_ga + state.run_s * k) / (k + 1)
    state.step = k + 1


def forward(a: torch.Tensor, b: torch.Tensor, p: MMTMParams, state: MMTMState, mode: int = MODE_NORMAL,
            avg: Optional[Sequence[torch.Tensor]] = None, gate_scale: float = 1.0):
    """One MMTM forward.  Mutates `state`.  Returns (A', B', aux)."""
    s_a, s_b, g_a, g_b, hs = gates(a, b, p, mode, avg)
    update_running(state, g_a)
    bc = (slice(None), slice(None)) + (None,) * (a.dim() - 2)
    if mode == MODE_CURATE_VISUAL:
        # :139-143 -- constant per-channel scale, already including this batch.
        ga_used = state.run_v.unsqueeze(0).expand_as(g_a)
        gb_used = g_b
    elif mode == MODE_CURATE_SKELETON:
        ga_used = g_a
        gb_used = state.run_s.unsqueeze(0).expand_as(g_b)
    else:
        ga_used, gb_used = g_a, g_b
    if gate_scale != 1.0:
        ga_used, gb_used = ga_used * gate_scale, gb_used * gate_scale
    a_out = a * ga_used[bc]
    b_out = b * gb_used[bc]
    aux = dict(sA=s_a, sB=s_b, gA=g_a, gB=g_b, H=hs)
    return a_out, b_out, aux


def forward_backward(a, b, p: MMTMParams, state: MMTMState, grad_a_out, grad_b_out, mode=MODE_NORMAL,
                     avg=None, gate_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """fp32 torch-autograd forward+backward; the CPU baseline `bench.py` times."""
    a = a.detach().clone().requires_grad_(True)
    b = b.detach().clone().requires_grad_(True)
    q = p.detach_clone(requires_grad=True)
    a_out, b_out, aux = forward(a, b, q, state, mode, avg, gate_scale)
    torch.autograd.backward([a_out, b_out], [grad_a_out, grad_b_out])
    z = lambda t: torch.zeros_like(t) if t.grad is None else t.grad
    return dict(A_out=a_out.detach(), B_out=b_out.detach(), dA=a.grad, dB=b.grad,
                dWsq=z(q.w_sq), dbsq=z(q.b_sq), dWv=z(q.w_v), dbv=z(q.b_v), dWs=z(q.w_s), dbs=z(q.b_s),
                has_grad=dict(w_v=q.w_v.grad is not None, w_s=q.w_s.grad is not None),
                sA=aux["sA"].detach(), sB=aux["sB"].detach(), gA=aux["gA"].detach(), gB=aux["gB"].detach())


# --------------------------------------------------------------------------------------
# float64 closed form (numpy).  Hand-derived backward; this is the set of formulas the
# CUDA backward kernels implement, stated once in the clearest possible way.
# --------------------------------------------------------------------------------------
def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def forward_backward_f64(a, b, w_sq, b_sq, w_v, b_v, w_s, b_s, grad_a_out, grad_b_out, run_v, run_s, step,
                         mode=MODE_NORMAL, avg=None, gate_scale=1.0):
    f = lambda t: np.asarray(t, dtype=np.float64)
    a, b, ga_o, gb_o = f(a), f(b), f(grad_a_out), f(grad_b_out)
    w_sq, b_sq, w_v, b_v, w_s, b_s = map(f, (w_sq, b_sq, w_v, b_v, w_s, b_s))
    n, c_v = a.shape[:2]
    c_s = b.shape[1]
    a3, b3 = a.reshape(n, c_v, -1), b.reshape(n, c_s, -1)
    hw_a, hw_b = a3.shape[2], b3.shape[2]
    s_a, s_b = a3.mean(-1), b3.mean(-1)
    if mode == MODE_XMODAL_OFF:
        z1 = np.concatenate([s_a, np.broadcast_to(f(avg[1]), (n, c_s))], 1)
        z2 = np.concatenate([np.broadcast_to(f(avg[0]), (n, c_v)), s_b], 1)
    else:
        z1 = z2 = np.concatenate([s_a, s_b], 1)
    pre1, pre2 = z1 @ w_sq.T + b_sq, z2 @ w_sq.T + b_sq
    h1, h2 = np.maximum(pre1, 0), np.maximum(pre2, 0)
    g_a, g_b = _sig(h1 @ w_v.T + b_v), _sig(h2 @ w_s.T + b_s)
    mean_ga = g_a.mean(0)
    run_v = (mean_ga + f(run_v) * step) / (step + 1)
    run_s = (mean_ga + f(run_s) * step) / (step + 1)
    live_a, live_b = mode != MODE_CURATE_VISUAL, mode != MODE_CURATE_SKELETON
    ga_used = g_a if live_a else np.broadcast_to(run_v, g_a.shape)
    gb_used = g_b if live_b else np.broadcast_to(run_s, g_b.shape)
    ga_used, gb_used = ga_used * gate_scale, gb_used * gate_scale
    a_out = a3 * ga_used[:, :, None]
    b_out = b3 * gb_used[:, :, None]
    ga3, gb3 = ga_o.reshape(n, c_v, -1), gb_o.reshape(n, c_s, -1)
    # d loss / d gate = sum_hw grad_out * input   (only where the gate is live)
    dg_a = (ga3 * a3).sum(-1) * gate_scale if live_a else np.zeros_like(g_a)
    dg_b = (gb3 * b3).sum(-1) * gate_scale if live_b else np.zeros_like(g_b)
    de_a = dg_a * g_a * (1 - g_a)
    de_b = dg_b * g_b * (1 - g_b)
    dh1 = (de_a @ w_v) * (pre1 > 0)
    dh2 = (de_b @ w_s) * (pre2 > 0)
    if mode == MODE_XMODAL_OFF:
        dz1, dz2 = dh1 @ w_sq, dh2 @ w_sq
        ds_a, ds_b = dz1[:, :c_v], dz2[:, c_v:]
        d_wsq = dh1.T @ z1 + dh2.T @ z2
        d_bsq = dh1.sum(0) + dh2.sum(0)
    else:
        dh = dh1 + dh2
        dz = dh @ w_sq
        ds_a, ds_b = dz[:, :c_v], dz[:, c_v:]
        d_wsq = dh.T @ z1
        d_bsq = dh.sum(0)
    d_a = ga3 * ga_used[:, :, None] + ds_a[:, :, None] / hw_a
    d_b = gb3 * gb_used[:, :, None] + ds_b[:, :, None] / hw_b
    return dict(A_out=a_out.reshape(a.shape), B_out=b_out.reshape(b.shape), dA=d_a.reshape(a.shape),
                dB=d_b.reshape(b.shape), dWsq=d_wsq, dbsq=d_bsq, dWv=de_a.T @ h1, dbv=de_a.sum(0),
                dWs=de_b.T @ h2, dbs=de_b.sum(0), sA=s_a, sB=s_b, gA=g_a, gB=g_b, run_v=run_v, run_s=run_s,
                step=step + 1)


# --------------------------------------------------------------------------------------
# Deterministic synthetic inputs shared by the golden generator and the tests.  numpy's
# legacy RandomState stream is stable across numpy versions, so fixtures do not need to
# store inputs.
# --------------------------------------------------------------------------------------
def synth_inputs(seed: int, n: int, c: int, h: int, w: Optional[int] = None):
    w = h if w is None else w
    rs = np.random.RandomState(seed)
    mk = lambda: torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32))
    return dict(A=mk(), B=mk(), gA=mk(), gB=mk())


def synth_params(seed: int, c_v: int, c_s: int, ratio=4) -> MMTMParams:
    """Uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default init, but drawn
    from numpy so the values do not depend on the torch version."""
    d = hidden_dim(c_v, c_s, ratio)
    rs = np.random.RandomState(seed)

    def u(shape, fan_in):
        k = 1.0 / np.sqrt(fan_in)
        return torch.from_numpy(rs.uniform(-k, k, size=shape).astype(np.float32))

    return MMTMParams(u((d, c_v + c_s), c_v + c_s), u((d,), c_v + c_s), u((c_v, d), d), u((c_v,), d),
                      u((c_s, d), d), u((c_s,), d))
