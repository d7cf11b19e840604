"""Drop-in mirror of the reference's `src/balanced_mmtm.py` on top of libgml_b200 (sm_100a).

Same public names, constructor/forward signatures, parameter names (checkpoint keys
`fc_squeeze|fc_visual|fc_skeleton.{weight,bias}`), return tuple and error behaviour as
`MMTM_mitigate` (reference src/balanced_mmtm.py:15-154) and `get_rescale_weights`
(:179-206).  The arithmetic runs in hand-written CUDA kernels behind the C ABI in
include/gml_b200.h; there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import os
import pickle
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import MMTMDims, MODE_CURATE_SKELETON, MODE_CURATE_VISUAL, MODE_NORMAL, MODE_XMODAL_OFF


def _mode_from_flags(curation_mode, caring_modality, turnoff_cross_modal_flow) -> int:
    """Reference control flow at balanced_mmtm.py:71-72,128-152 folded into one id."""
    if turnoff_cross_modal_flow:
        if curation_mode:
            raise NotImplementedError("curation_mode with turnoff_cross_modal_flow is never used by the reference")
        return MODE_XMODAL_OFF
    if not curation_mode:
        return MODE_NORMAL
    if caring_modality == 0:
        return MODE_CURATE_VISUAL
    if caring_modality == 1:
        return MODE_CURATE_SKELETON
    # reference: neither branch reshapes the [N, C] gates -> broadcasting against NCHW fails
    raise RuntimeError("curation_mode=True requires caring_modality in {0, 1} (got %r)" % (caring_modality,))


def _dims(a, b, d) -> MMTMDims:
    n, c_v = a.shape[0], a.shape[1]
    c_s = b.shape[1]
    hw_v = a[0, 0].numel() if n else int(np.prod(a.shape[2:]))
    hw_s = b[0, 0].numel() if n else int(np.prod(b.shape[2:]))
    return MMTMDims(n, c_v, c_s, hw_v, hw_s, d)


class _MMTMFunction(torch.autograd.Function):
    """autograd node = one gml_mmtm_fwd launch sequence / one gml_mmtm_bwd."""

    @staticmethod
    def forward(ctx, a, b, w_sq, b_sq, w_v, b_v, w_s, b_s, run_v, run_s, step, mode, gate_scale, m_a, m_b, flags,
                dist_group):
        lib = _lib.load()
        dev = a.device
        d = w_sq.shape[0]
        dims = _dims(a, b, d)
        n, c_v, c_s = dims.n, dims.c_v, dims.c_s
        zrows = 2 * n if mode == MODE_XMODAL_OFF else n
        f32 = dict(dtype=torch.float32, device=dev)
        a_out, b_out = torch.empty_like(a), torch.empty_like(b)
        z = torch.empty((zrows, c_v + c_s), **f32)
        h = torch.empty((zrows, d), **f32)
        g_a, g_b = torch.empty((n, c_v), **f32), torch.empty((n, c_s), **f32)
        gate_sum = torch.empty((c_v + 1,), **f32)  # [sum_n g_a | n] so one all-reduce carries both
        st = _lib.current_stream(dev)
        P = _lib.ptr
        ws_bytes = lib.gml_mmtm_fwd_workspace_bytes(dims)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        curate = mode in (MODE_CURATE_VISUAL, MODE_CURATE_SKELETON)
        world = torch.distributed.get_world_size(dist_group) if dist_group is not None else 1
        with torch.cuda.device(dev):
            if world == 1:
                _lib.check(lib.gml_mmtm_fwd(P(a), P(b), P(a_out), P(b_out), P(w_sq), P(b_sq), P(w_v), P(b_v), P(w_s),
                                            P(b_s), P(z), P(h), P(g_a), P(g_b), P(gate_sum), P(run_v), P(run_s), step,
                                            P(m_a), P(m_b), P(ws), ws_bytes, dims, mode, gate_scale, flags, st),
                           "gml_mmtm_fwd")
            else:
                # data parallel: the running mean is over the GLOBAL batch (one tiny all-reduce of
                # per-rank gate sums).  In the curation modes the substituted scale
                # depends on it, so the gating pass has to wait for the collective.
                gate_sum[c_v] = float(n)
                if curate:
                    _lib.check(lib.gml_mmtm_gates(P(a), P(b), P(w_sq), P(b_sq), P(w_v), P(b_v), P(w_s), P(b_s), P(z),
                                                  P(h), P(g_a), P(g_b), P(gate_sum), P(m_a), P(m_b), P(ws), ws_bytes,
                                                  dims, mode, st), "gml_mmtm_gates")
                else:
                    _lib.check(lib.gml_mmtm_fwd(P(a), P(b), P(a_out), P(b_out), P(w_sq), P(b_sq), P(w_v), P(b_v),
                                                P(w_s), P(b_s), P(z), P(h), P(g_a), P(g_b), P(gate_sum), P(run_v),
                                                P(run_s), step, P(m_a), P(m_b), P(ws), ws_bytes, dims, mode, gate_scale,
                                                flags | _lib.F_NO_RUNNING_UPDATE, st), "gml_mmtm_fwd")
                torch.distributed.all_reduce(gate_sum, group=dist_group)
                # keep the count on device: divide by it inside a tiny torch op instead of syncing
                _MMTMFunction._running_update_dp(run_v, run_s, gate_sum[:c_v], gate_sum[c_v], step)
                if curate:
                    _lib.check(lib.gml_mmtm_apply(P(a), P(b), P(a_out), P(b_out), P(g_a), P(g_b), P(run_v), P(run_s),
                                                  dims, mode, gate_scale, st), "gml_mmtm_apply")
        ctx.mode, ctx.gate_scale, ctx.dims, ctx.flags = mode, gate_scale, dims, flags
        run_saved_v = run_v.clone() if mode == MODE_CURATE_VISUAL else None
        run_saved_s = run_s.clone() if mode == MODE_CURATE_SKELETON else None
        ctx.save_for_backward(a, b, w_sq, w_v, w_s, z, h, g_a, g_b, run_saved_v, run_saved_s)
        ctx.mark_non_differentiable(z, g_a, g_b)
        return a_out, b_out, z, g_a, g_b

    @staticmethod
    def _running_update_dp(run_v, run_s, gsum, n_total, step):
        # (mean + run * step) / (step + 1) with the all-reduced sum and count, both on device
        mean = gsum / n_total
        run_v.mul_(float(step)).add_(mean).div_(float(step + 1))
        run_s.mul_(float(step)).add_(mean).div_(float(step + 1))

    @staticmethod
    def backward(ctx, go_a, go_b, _gz, _gga, _ggb):
        lib = _lib.load()
        a, b, w_sq, w_v, w_s, z, h, g_a, g_b, run_v, run_s = ctx.saved_tensors
        dev = a.device
        dims, mode = ctx.dims, ctx.mode
        go_a = torch.zeros_like(a) if go_a is None else go_a.contiguous()
        go_b = torch.zeros_like(b) if go_b is None else go_b.contiguous()
        need = ctx.needs_input_grad
        d_a, d_b = torch.empty_like(a), torch.empty_like(b)
        live_a, live_b = mode != MODE_CURATE_VISUAL, mode != MODE_CURATE_SKELETON
        mk = lambda t, on: torch.empty_like(t) if on else None
        d_w_sq, d_b_sq = mk(w_sq, need[2]), (torch.empty(w_sq.shape[0], device=dev) if need[3] else None)
        d_w_v, d_b_v = mk(w_v, need[4] and live_a), (torch.empty(w_v.shape[0], device=dev) if need[5] and live_a else None)
        d_w_s, d_b_s = mk(w_s, need[6] and live_b), (torch.empty(w_s.shape[0], device=dev) if need[7] and live_b else None)
        ws_bytes = lib.gml_mmtm_bwd_workspace_bytes(dims)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        P = _lib.ptr
        with torch.cuda.device(dev):
            _lib.check(lib.gml_mmtm_bwd(P(go_a), P(go_b), P(a), P(b), P(w_sq), P(w_v), P(w_s), P(z), P(h), P(g_a),
                                        P(g_b), P(run_v), P(run_s), None, None, P(d_a), P(d_b), P(d_w_sq), P(d_b_sq),
                                        P(d_w_v), P(d_b_v), P(d_w_s), P(d_b_s), P(ws), ws_bytes, dims, mode,
                                        ctx.gate_scale, ctx.flags, _lib.current_stream(dev)), "gml_mmtm_bwd")
        # a substituted side's excitation FC gets no gradient (None, like the reference's autograd)
        return (d_a if need[0] else None, d_b if need[1] else None, d_w_sq, d_b_sq, d_w_v, d_b_v, d_w_s, d_b_s,
                None, None, None, None, None, None, None, None, None)


class MMTM_mitigate(nn.Module):
    """CUDA-kernel MMTM fusion block; interface of reference src/balanced_mmtm.py:15-154.

    Differences that are deliberate and documented:
      * running statistics follow the module's device instead of being pinned to
        ``cuda:{device}`` at construction (the reference's ctor needs a GPU, :30-31);
      * ``SEonly`` / ``shareweight`` are accepted but only the default (False, False)
        path exists -- the reference never instantiates the others (src/model.py:58-60);
      * ``gate_scale`` (default 1.0 = reference's plain sigmoid) is an extra keyword;
      * under ``torch.distributed`` with ``sync_running_stats=True`` the running gate mean
        is the global-batch mean (one small all-reduce per forward).
    """

    def __init__(self, dim_visual, dim_skeleton, ratio, device=0, SEonly=False, shareweight=False, gate_scale=1.0,
                 sync_running_stats=True, kernel_flags=0):
        super().__init__()
        if SEonly or shareweight:
            raise NotImplementedError("SEonly / shareweight are dead code in the reference (never instantiated)")
        dim = dim_visual + dim_skeleton
        dim_out = int(2 * dim / ratio)
        self.SEonly, self.shareweight = SEonly, shareweight
        self.dim_visual, self.dim_skeleton = dim_visual, dim_skeleton
        # non-persistent running statistics, plain attributes like the reference (:30-32):
        # never part of state_dict, reset only by re-constructing the module.
        self.running_avg_weight_visual = torch.zeros(dim_visual)
        self.running_avg_weight_skeleton = torch.zeros(dim_visual)
        self.step = 0
        # construction order = reference order (:37-45) so the global RNG stream matches
        self.fc_squeeze = nn.Linear(dim, dim_out)
        self.fc_visual = nn.Linear(dim_out, dim_visual)
        self.fc_skeleton = nn.Linear(dim_out, dim_skeleton)
        self.relu = nn.ReLU()
        self.sigmoid = nn.Sigmoid()
        self.gate_scale = float(gate_scale)
        self.sync_running_stats = sync_running_stats
        self.kernel_flags = int(kernel_flags)
        self.process_group = None  # None = default group when torch.distributed is initialised

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.running_avg_weight_visual = fn(self.running_avg_weight_visual)
        self.running_avg_weight_skeleton = fn(self.running_avg_weight_skeleton)
        return self

    def _dist_group(self):
        if not self.sync_running_stats:
            return None
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            return self.process_group if self.process_group is not None else dist.group.WORLD
        return None

    def forward(self, visual, skeleton, return_scale=False, return_squeezed_mps=False,
                turnoff_cross_modal_flow=False, average_squeezemaps=None, curation_mode=False, caring_modality=0):
        mode = _mode_from_flags(curation_mode, caring_modality, turnoff_cross_modal_flow)
        _lib.require_cuda(visual, skeleton, self.fc_squeeze.weight)
        if visual.dtype != torch.float32 or skeleton.dtype != torch.float32:
            raise NotImplementedError("fp32 feature maps only (the reference's path is fp32)")
        if visual.shape[0] != skeleton.shape[0]:
            raise ValueError("batch sizes differ: %s vs %s" % (tuple(visual.shape), tuple(skeleton.shape)))
        if visual.shape[1] != self.dim_visual or skeleton.shape[1] != self.dim_skeleton:
            raise ValueError("channel mismatch")
        if return_squeezed_mps and mode == MODE_XMODAL_OFF:
            # the reference never defines squeeze_array on this branch (:72-91,123-124)
            raise UnboundLocalError("local variable 'squeeze_array' referenced before assignment")
        dev = visual.device
        if self.running_avg_weight_visual.device != dev:
            self.running_avg_weight_visual = self.running_avg_weight_visual.to(dev)
            self.running_avg_weight_skeleton = self.running_avg_weight_skeleton.to(dev)
        m_a = m_b = None
        if mode == MODE_XMODAL_OFF:
            m_a = average_squeezemaps[0].detach().to(device=dev, dtype=torch.float32).contiguous()
            m_b = average_squeezemaps[1].detach().to(device=dev, dtype=torch.float32).contiguous()
        if visual.shape[0] == 0:
            # empty batch: nothing to launch (the reference would produce NaN running means here).  Under data
            # parallelism the other ranks are inside the gate-sum all-reduce: join it with zeros and a count of 0
            grp = self._dist_group()
            if grp is not None:
                gs = torch.zeros(self.dim_visual + 1, dtype=torch.float32, device=dev)
                torch.distributed.all_reduce(gs, group=grp)
                _MMTMFunction._running_update_dp(self.running_avg_weight_visual, self.running_avg_weight_skeleton,
                                                 gs[:self.dim_visual], gs[self.dim_visual], self.step)
                self.step += 1
            return visual.clone(), skeleton.clone(), None, None
        a, b = visual.contiguous(), skeleton.contiguous()
        # the reference rebinds running_avg_* to fresh tensors every call (:113-114); keep that
        # aliasing behaviour (a caller holding the old tensor does not see it change)
        run_v = self.running_avg_weight_visual.clone()
        run_s = self.running_avg_weight_skeleton.clone()
        a_out, b_out, z, g_a, g_b = _MMTMFunction.apply(
            a, b, self.fc_squeeze.weight, self.fc_squeeze.bias, self.fc_visual.weight, self.fc_visual.bias,
            self.fc_skeleton.weight, self.fc_skeleton.bias, run_v, run_s, self.step, mode, self.gate_scale, m_a, m_b,
            self.kernel_flags, self._dist_group())
        self.running_avg_weight_visual, self.running_avg_weight_skeleton = run_v, run_s
        self.step += 1
        scales = [g_a.cpu(), g_b.cpu()] if return_scale else None
        squeeze_array = None
        if return_squeezed_mps:
            squeeze_array = [z[:, :self.dim_visual].cpu(), z[:, self.dim_visual:].cpu()]
        self.last_squeeze = z  # device-side handle for SqueezeMeanRecorder (no host copy)
        return a_out, b_out, scales, squeeze_array


# ------------------------------------------------------------------------------------------
# conditional utilization inputs
# ------------------------------------------------------------------------------------------
def get_mmtm_outputs(eval_save_path, mmtm_recorded, key):
    """File-format reader, reference src/balanced_mmtm.py:157-176 (host logic, no kernels)."""
    with open(os.path.join(eval_save_path, "history.pickle"), "rb") as f:
        his_epo = pickle.load(f)
    order = np.argsort(his_epo["test_indices"][0])
    data = []
    for batch in his_epo[key][0]:
        assert mmtm_recorded == len(batch)
        for mmtmid, views in enumerate(batch):
            if len(data) < mmtmid + 1:
                data.append({})
            for i, view in enumerate(views):
                arr = view.detach().cpu().numpy() if torch.is_tensor(view) else np.asarray(view)
                data[mmtmid].setdefault("view_%d" % i, []).append(arr)
    for blk in data:
        for k in blk:
            blk[k] = np.concatenate(blk[k])[order]
    return data


def get_rescale_weights(eval_save_path, training_save_path, key="test_squeezedmaps_array_list", validation=False,
                        starting_mmtmindice=1, mmtmpositions=4, device=None):
    """Dataset-mean squeezes per MMTM block from the recorded history pickles; same
    signature and return layout as reference src/balanced_mmtm.py:179-206:
    ``[None] * starting_mmtmindice + [[mean_view0, mean_view1], ...]``."""
    data = get_mmtm_outputs(eval_save_path, mmtmpositions - starting_mmtmindice, key)
    with open(os.path.join(training_save_path, "history.pickle"), "rb") as f:
        his_ori = pickle.load(f)
    selected = his_ori["val_indices"][0] if validation else his_ori["train_indices"][0]
    out = []
    for pos in range(mmtmpositions):
        if pos < starting_mmtmindice:
            out.append(None)
            continue
        blk = data[pos - starting_mmtmindice]
        weights = [blk[k][selected].mean(0) for k in sorted(blk.keys())]
        if device is not None:
            weights = [torch.from_numpy(w).to(device) for w in weights]
        out.append(weights)
    return out


class SqueezeMeanRecorder:
    """On-device replacement for the recording.gin -> history.pickle -> get_rescale_weights
    round trip (reference src/framework.py:160-161,244-247; src/balanced_mmtm.py:186-201).

    Call `update(blocks, indices)` after every forward of the recording pass; it folds the
    squeezes of the SELECTED dataset indices into fp64 device sums with
    gml_squeeze_accumulate (no per-batch D2H copy).  `result()` all-reduces the sums across
    data-parallel ranks and returns the structure get_rescale_weights returns.
    """

    def __init__(self, mmtm_blocks: Sequence[MMTM_mitigate], selected_indices=None, starting_mmtmindice=1):
        self.blocks = list(mmtm_blocks)
        self.starting = starting_mmtmindice
        self.selected = None if selected_indices is None else torch.as_tensor(np.asarray(selected_indices)).long()
        self._sums = None
        self._count = None
        self._lookup = None

    def _lazy(self, dev):
        if self._sums is None:
            self._sums = [torch.zeros(b.dim_visual + b.dim_skeleton, dtype=torch.float64, device=dev)
                          for b in self.blocks]
            self._count = torch.zeros(len(self.blocks), dtype=torch.int64, device=dev)
            if self.selected is not None:
                size = int(self.selected.max().item()) + 1 if self.selected.numel() else 0
                self._lookup = torch.zeros(size, dtype=torch.uint8, device=dev)
                self._lookup[self.selected.to(dev)] = 1

    def update(self, indices=None):
        lib = _lib.load()
        z0 = self.blocks[0].last_squeeze
        dev = z0.device
        self._lazy(dev)
        sel = None
        if self._lookup is not None:
            idx = torch.as_tensor(indices).to(dev).long()
            inside = idx < self._lookup.numel()
            sel = torch.zeros(idx.numel(), dtype=torch.uint8, device=dev)
            sel[inside] = self._lookup[idx[inside]]
        with torch.cuda.device(dev):
            for i, blk in enumerate(self.blocks):
                z = blk.last_squeeze
                _lib.check(lib.gml_squeeze_accumulate(z.data_ptr(), _lib.ptr(sel), z.shape[0], z.shape[1],
                                                      self._sums[i].data_ptr(), self._count[i:].data_ptr(),
                                                      _lib.current_stream(dev)), "gml_squeeze_accumulate")

    def result(self, device=None, group=None):
        dist = torch.distributed
        sums = [s.clone() for s in self._sums]
        count = self._count.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for s in sums:
                dist.all_reduce(s, group=group)
            dist.all_reduce(count, group=group)
        out = [None] * self.starting
        for i, blk in enumerate(self.blocks):
            mean = (sums[i] / count[i].double()).float()
            views = [mean[:blk.dim_visual].contiguous(), mean[blk.dim_visual:].contiguous()]
            if device is not None:
                views = [v.to(device) for v in views]
            out.append(views)
        return out
