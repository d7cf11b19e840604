"""Rebalancing controllers: mirrors of the reference's `Bias_Mitigation_Strong` and
`Bias_Mitigation_Random` (src/callbacks.py:173-302) with the conditional-learning-speed
statistic computed by ONE multi-tensor CUDA launch and ONE 64-byte read-back per step
(gml_multi_tensor_sqnorm) instead of 2 reductions + 2 `.item()` syncs per parameter.

The hook surface is the reference's Keras-style `Callback` (src/callbacks.py:97-170), so
instances can be dropped into the reference's own `CallbackList` / `training_loop`.
"""
from __future__ import annotations

import ctypes
import random
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import gin_lite


class Callback(object):
    """Hook surface of reference src/callbacks.py:97-170 (setters + no-op hooks)."""

    def __init__(self):
        pass

    def set_config(self, config): self.config = config
    def set_meta_data(self, meta_data): self.meta_data = meta_data
    def set_save_path(self, save_path): self.save_path = save_path
    def set_optimizer(self, optimizer): self.optimizer = optimizer

    def set_model(self, model, ignore=True):
        if ignore:
            return
        self.model = model

    def set_model_pytoune(self, model_pytoune): self.model_pytoune = model_pytoune
    def set_params(self, params): self.params = params
    def set_dataloader(self, data): self.data = data
    def get_dataloader(self): return self.data
    def get_config(self): return self.config
    def get_meta_data(self): return self.meta_data
    def get_optimizer(self): return self.optimizer
    def get_params(self): return self.params
    def get_model(self): return self.model
    def get_save_path(self): return self.save_path
    def on_epoch_begin(self, epoch, logs): pass
    def on_epoch_end(self, epoch, logs): pass
    def on_batch_begin(self, batch, logs): pass
    def on_batch_end(self, batch, logs): pass
    def on_forward_begin(self, batch, data): pass
    def on_backward_end(self, batch): pass
    def on_train_begin(self, logs): pass
    def on_train_end(self, logs): pass
    def on_val_batch_end(self, batch, logs): pass


def bucket_mask(name: str, branchnames: Sequence[str], mmtmnames: Sequence[str]) -> int:
    """Bucket bits for one parameter name; string rules of reference src/callbacks.py:207-223."""
    mask = 0
    if "mmtm" in name:
        bits = (_lib.BUCKET_BYPASS0, _lib.BUCKET_BYPASS1)
        tagged = [i for i, tag in enumerate(mmtmnames) if tag in name]
        for i in (tagged or range(len(mmtmnames))):  # untagged (fc_squeeze) -> every modality
            mask |= bits[i]
    else:
        bits = (_lib.BUCKET_MAIN0, _lib.BUCKET_MAIN1)
        for i, tag in enumerate(branchnames):
            if tag in name:
                mask |= bits[i]
    return mask


class MultiTensorSqnorm:
    """Host-side table for gml_multi_tensor_sqnorm over a model's parameters and gradients.

    The (pointer, numel, bucket, kind) table is rebuilt only when a data pointer changed
    (optimizer.zero_grad(set_to_none=True) re-allocates gradients; the caching allocator
    usually hands back the same addresses).
    """

    def __init__(self, named_parameters, branchnames, mmtmnames):
        self.entries = [(n, p) for n, p in named_parameters]
        self.masks = [bucket_mask(n, branchnames, mmtmnames) for n, _ in self.entries]
        self._key = None
        self._out = None

    def _build(self, dev):
        n = 2 * len(self.entries)
        ptrs = (ctypes.c_void_p * n)()
        numel = (ctypes.c_int64 * n)()
        masks = (ctypes.c_int32 * n)()
        kinds = (ctypes.c_int32 * n)()
        for i, ((name, p), m) in enumerate(zip(self.entries, self.masks)):
            g = p.grad
            if g is None:
                raise RuntimeError("parameter %s has no gradient (compute_BDR is only valid after a normal-mode "
                                   "backward; reference src/callbacks.py:204 would raise too)" % name)
            if p.dtype != torch.float32 or g.dtype != torch.float32:
                raise NotImplementedError("fp32 parameters only")
            if not p.is_contiguous() or not g.is_contiguous():
                raise NotImplementedError("contiguous parameters/gradients only")
            ptrs[2 * i], numel[2 * i], masks[2 * i], kinds[2 * i] = p.data_ptr(), p.numel(), m, 0
            ptrs[2 * i + 1], numel[2 * i + 1], masks[2 * i + 1], kinds[2 * i + 1] = g.data_ptr(), g.numel(), m, 1
        lib = _lib.load()
        ws_bytes = lib.gml_sqnorm_workspace_bytes(numel, n)
        self._table = (ptrs, numel, masks, kinds, n)
        self._ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        self._ws_bytes = ws_bytes
        if self._out is None:
            self._out = torch.empty(8, dtype=torch.float64, device=dev)
            self._host = torch.empty(8, dtype=torch.float64).pin_memory()

    def measure(self) -> Dict[str, List[float]]:
        lib = _lib.load()
        p0 = self.entries[0][1]
        _lib.require_cuda(p0)
        dev = p0.device
        key = tuple((p.data_ptr(), -1 if p.grad is None else p.grad.data_ptr()) for _, p in self.entries)
        if key != self._key:
            self._build(dev)
            self._key = key
        ptrs, numel, masks, kinds, n = self._table
        with torch.cuda.device(dev):
            _lib.check(lib.gml_multi_tensor_sqnorm(ptrs, numel, masks, kinds, n, self._out.data_ptr(), None,
                                                   self._ws.data_ptr(), self._ws_bytes, _lib.current_stream(dev)),
                       "gml_multi_tensor_sqnorm")
        self._host.copy_(self._out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()  # the single host sync of the statistic
        v = self._host.tolist()
        return dict(wn_main=v[0:2], wn_bypass=v[2:4], gn_main=v[4:6], gn_bypass=v[6:8])


class Bias_Mitigation_Strong(Callback):
    """Guided rebalancing controller; state machine of reference src/callbacks.py:173-267."""

    def __init__(self, epsilon, curation_windowsize, branchnames, starting_epoch=2, MMTMnames=['visual', 'skeleton']):
        self.epsilon = epsilon
        self.branchnames = branchnames
        self.MMTMnames = MMTMnames
        self.curation_windowsize = curation_windowsize
        self.starting_epoch = starting_epoch
        self._sqnorm = None
        super(Bias_Mitigation_Strong, self).__init__()

    def on_train_begin(self, logs):
        self.M_bypass_modal_0 = 0
        self.M_bypass_modal_1 = 0
        self.M_main_modal_0 = 0
        self.M_main_modal_1 = 0
        self.model_pytoune.curation_mode = False
        self.model_pytoune.caring_modality = None
        self.unlock = False

    def measure_sqnorms(self) -> Dict[str, List[float]]:
        """The 8 bucket sums for the current parameters/gradients (one CUDA launch)."""
        if self._sqnorm is None:
            self._sqnorm = MultiTensorSqnorm(self.model.named_parameters(), self.branchnames, self.MMTMnames)
        return self._sqnorm.measure()

    def compute_BDR(self):
        b = self.measure_sqnorms()
        # accumulated since train begin, never reset; double arithmetic (callbacks.py:225-233)
        self.M_bypass_modal_0 += b["gn_bypass"][0] / b["wn_bypass"][0]
        self.M_bypass_modal_1 += b["gn_bypass"][1] / b["wn_bypass"][1]
        self.M_main_modal_0 += b["gn_main"][0] / b["wn_main"][0]
        self.M_main_modal_1 += b["gn_main"][1] / b["wn_main"][1]
        BDR_0 = np.log10(self.M_bypass_modal_0 / self.M_main_modal_0)
        BDR_1 = np.log10(self.M_bypass_modal_1 / self.M_main_modal_1)
        return BDR_0 - BDR_1

    def on_batch_end(self, batch, logs):
        logs['curation_mode'] = float(self.model_pytoune.curation_mode)
        logs['caring_modality'] = self.model_pytoune.caring_modality
        logs['d_BDR'] = self.d_BDR

    def on_backward_end(self, batch):
        """Decision for the NEXT step (it runs before optimizer.step, src/framework.py:313-315):
        locked -> keep accumulating the statistic, normal mode; unlocked and normal -> measure and, when the two
        conditional learning speeds differ by more than epsilon, open a window that cares for the slower modality;
        inside a window -> count steps only (the statistic is neither measured nor accumulated there)."""
        mp = self.model_pytoune
        if self.unlock and mp.curation_mode:
            self.curation_step += 1
            if self.curation_step == self.curation_windowsize:
                mp.curation_mode = False
            return
        self.d_BDR = self.compute_BDR()
        imbalanced = self.unlock and abs(self.d_BDR) > self.epsilon
        mp.curation_mode = bool(imbalanced)
        if imbalanced:
            self.curation_step = 0
            direction = np.sign(self.d_BDR)
            if direction > 0:
                mp.caring_modality = 0
            elif direction < 0:
                mp.caring_modality = 1
            # (a NaN statistic leaves caring_modality as it was, like the reference's two-way branch)
        else:
            mp.caring_modality = 0

    def on_epoch_begin(self, epoch, logs):
        if epoch >= self.starting_epoch:
            self.unlock = True


class Bias_Mitigation_Random(Callback):
    """Random rebalancing controller (reference src/callbacks.py:269-302).  Uses the global
    `random` module like the reference; under data parallelism every rank must seed it
    identically (greedy_multimodal_learning_b200.dist.seed_everything does)."""

    def on_train_begin(self, logs):
        self.model_pytoune.curation_mode = False
        self.model_pytoune.caring_modality = None
        self.unlock = False
        self.starting_epoch = 2

    def on_batch_end(self, batch, logs):
        logs['curation_mode'] = float(self.model_pytoune.curation_mode)
        logs['caring_modality'] = self.model_pytoune.caring_modality

    def on_backward_end(self, batch):
        mp = self.model_pytoune
        if self.unlock:
            mode = random.choice([0, 1, 2])
            if mode == 0:
                mp.curation_mode, mp.caring_modality = False, 0
            elif mode == 1:
                mp.curation_mode, mp.caring_modality = True, 1
            else:
                mp.curation_mode, mp.caring_modality = True, 0
        else:
            mp.curation_mode, mp.caring_modality = False, 0

    def on_epoch_begin(self, epoch, logs):
        if epoch >= self.starting_epoch:
            self.unlock = True


# ---------------------------------------------------------------------------------------------
# Auxiliary callbacks named by the reference's gin files (`train.callbacks=[...]`,
# configs/training*.gin) and by its default callback set (src/training_loop.py:27-51).  Host
# logic only; kept so that the five config files run unchanged.
# ---------------------------------------------------------------------------------------------
class CompletedStopping(Callback):
    """Stop after `patience` epochs whose `monitor` equals 100 (reference src/callbacks.py:306-333;
    the counter never resets between non-perfect epochs)."""

    def __init__(self, *, monitor='acc', patience=5, verbose=True):
        super().__init__()
        self.monitor, self.patience, self.verbose = monitor, patience, verbose
        self.stopped_epoch = 0

    def on_train_begin(self, logs):
        self.stopped_epoch, self.counter = 0, 0

    def on_epoch_end(self, epoch, logs):
        self.counter += int(logs[self.monitor] == 100)
        if self.counter >= self.patience:
            self.stopped_epoch = epoch
            self.model_pytoune.stop_training = True

    def on_train_end(self, logs):
        if self.stopped_epoch > 0 and self.verbose:
            print('Epoch %05d: completed stopping' % (self.stopped_epoch + 1))


class ReduceLROnPlateau_PyTorch(Callback):
    """torch ReduceLROnPlateau on an epoch log key (reference src/callbacks.py:336-351; the
    `verbose=` argument it passes no longer exists in torch 2.11 and is dropped)."""

    def __init__(self, metric, factor=0.3, patience=10):
        self.metric, self.factor, self.patience = metric, factor, patience

    def on_train_begin(self, logs):
        optimizer = getattr(self.optimizer, "optimizer", self.optimizer)  # unwrap dist.DPOptimizer
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(
            optimizer, mode='min', factor=self.factor, patience=self.patience, threshold=0.001,
            threshold_mode='rel', cooldown=0, min_lr=1e-6, eps=1e-08)

    def on_epoch_end(self, epoch, logs):
        self.scheduler.step(logs[self.metric])


class LambdaCallback(Callback):
    """Callback from plain functions (reference src/callbacks.py:354-386)."""

    _HOOKS = ('on_epoch_begin', 'on_epoch_end', 'on_batch_begin', 'on_batch_end', 'on_train_begin', 'on_train_end')

    def __init__(self, **hooks):
        super().__init__()
        for name, fn in hooks.items():
            if name not in self._HOOKS:
                raise TypeError("LambdaCallback: unknown hook %r" % name)
            if fn is not None:
                setattr(self, name, fn)


def save_weights(model, optimizer, filename):
    """`{'model': state_dict, 'optimizer': state_dict}` -- the checkpoint layout the reference
    writes and reloads (src/utils.py:103-111, src/training_loop.py:78-83)."""
    torch.save({'model': model.state_dict(), 'optimizer': optimizer.state_dict()}, filename)


class ModelCheckpoint(Callback):
    """Save on improvement of `monitor` (reference src/callbacks.py:389-451)."""

    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, mode='auto', period=1):
        super().__init__()
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.period = save_best_only, period
        self.epochs_since_last_save = 0
        if mode not in ('min', 'max'):
            mode = 'max' if ('acc' in monitor or monitor.startswith('fmeasure')) else 'min'
        self.monitor_op = np.greater if mode == 'max' else np.less
        self.best = -np.inf if mode == 'max' else np.inf

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        if not self.save_best_only:
            # the reference only saves here when verbose > 0 (src/callbacks.py:448-451); kept
            if self.verbose > 0:
                print('Epoch %05d: saving model to %s' % (epoch, self.filepath))
                save_weights(self.model, self.optimizer, self.filepath)
            return
        current = logs.get(self.monitor)
        if current is None:
            return
        if self.monitor_op(current, self.best):
            if self.verbose > 0:
                print('Epoch %05d: %s improved from %0.5f to %0.5f, saving model to %s'
                      % (epoch, self.monitor, self.best, current, self.filepath))
            self.best = current
            save_weights(self.model, self.optimizer, self.filepath)
        elif self.verbose > 0:
            print('Epoch %05d: %s did not improve' % (epoch, self.monitor))


class ProgressionCallback(Callback):
    """One status line per epoch (the reference redraws a line per batch, src/callbacks.py:454-517)."""

    def __init__(self, other_metrics=('average_iol_current_epoch', 'average_iol')):
        self.other_metrics = list(other_metrics)

    def on_train_begin(self, logs):
        self.metrics = ['loss'] + list(self.model_pytoune.metrics_names)
        self.epochs = self.params['epochs']

    def on_epoch_end(self, epoch, logs):
        keys = self.metrics + ['val_' + k for k in self.metrics] + self.other_metrics
        shown = ', '.join('%s: %f' % (k, logs[k]) for k in keys
                          if isinstance(logs.get(k), (int, float, np.floating, np.integer)))
        print("Epoch %d/%d %.2fs: %s" % (epoch, self.epochs, logs.get('time', 0.0), shown))


# gin names of the reference (`@gin.configurable` in src/callbacks.py)
for _cls in (Bias_Mitigation_Strong, Bias_Mitigation_Random, CompletedStopping, ReduceLROnPlateau_PyTorch,
             ProgressionCallback):
    gin_lite.configurable(_cls)
del _cls
