"""Step engine around the hot path: mirror of the reference's `Model_` / `StepIterator`
(src/framework.py:36-345) and of `blend_loss` / `acc` (train.py:23-40).

Observable behaviour kept: hook order (on_batch_begin -> on_forward_begin -> forward ->
backward -> on_backward_end -> optimizer.step -> on_batch_end), size-weighted epoch means,
`curation_mode` / `caring_modality` attributes read by every forward (train AND the
epoch-end val/test passes), history keys (`loss`, `acc`, `acc_modal_i`, `val_*`, `test_*`,
`*_indices`, extra lists).

What changed: the three `acc` calls + `loss.item()` cost the reference four
host syncs per batch (framework.py:154-156,317); here argmax/equality counts come from one
kernel (gml_accuracy_counts) and travel to the host together with the loss in a single
read-back.  Counts are integers -> accuracies are bit-exact.
"""
from __future__ import annotations

import itertools
import math
import timeit
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .callbacks import Callback


def blend_loss(y_hat, y):
    """Sum over views of mean cross-entropy (reference train.py:23-29)."""
    loss_func = torch.nn.CrossEntropyLoss()
    return sum(loss_func(y_pred, y) for y_pred in y_hat)


def acc(y_pred, y_true):
    """Accuracy in percent, reference train.py:32-40 (incl. the len(y_true)==2 quirk)."""
    if isinstance(y_pred, list):
        y_pred = torch.mean(torch.stack([out.data for out in y_pred], 0), 0)
    _, y_pred = y_pred.max(1)
    if len(y_true) == 2:
        acc_pred = (y_pred == y_true[0]).float().mean()
    else:
        acc_pred = (y_pred == y_true).float().mean()
    return acc_pred * 100


def percent_from_count(k: int, n: int) -> float:
    """float32 `count / n * 100`, the value `float(acc(...))` yields on the reference's CPU path."""
    return float(np.float32(k) / np.float32(n) * np.float32(100))


class CallbackList:
    """reference src/callbacks.py:28-95."""

    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])

    def append(self, cb): self.callbacks.append(cb)
    def __iter__(self): return iter(self.callbacks)

    def _each(self, name, *args):
        for cb in self.callbacks:
            getattr(cb, name)(*args)

    def set_params(self, params): self._each("set_params", params)
    def set_model(self, model): self._each("set_model", model)
    def set_model_pytoune(self, mp): self._each("set_model_pytoune", mp)
    def on_epoch_begin(self, epoch, logs=None): self._each("on_epoch_begin", epoch, logs or {})
    def on_epoch_end(self, epoch, logs=None): self._each("on_epoch_end", epoch, logs or {})
    def on_batch_begin(self, batch, logs=None): self._each("on_batch_begin", batch, logs or {})
    def on_batch_end(self, batch, logs=None): self._each("on_batch_end", batch, logs or {})
    def on_forward_begin(self, batch, data): self._each("on_forward_begin", batch, data)
    def on_backward_end(self, batch): self._each("on_backward_end", batch)
    def on_train_begin(self, logs=None): self._each("on_train_begin", logs or {})
    def on_train_end(self, logs=None): self._each("on_train_end", logs or {})


def _cycle(iterable):
    while True:
        for x in iterable:
            yield x


class StepIterator:
    """reference src/framework.py:36-122 (accumulators + per-batch hook calls)."""

    default_fields = ('indices', 'loss', 'metrics', 'viewwises_metrics', 'number', 'size')

    def __init__(self, generator, steps_per_epoch, callback, metrics_names, nummodalities):
        self.generator, self.steps_per_epoch, self.callback = generator, steps_per_epoch, callback
        self.metrics_names, self.nummodalities = metrics_names, nummodalities
        self.losses_sum = 0.
        self.metrics_sum = np.zeros(len(metrics_names))
        self.metrics_permodal_sum = np.zeros((nummodalities, len(metrics_names)))
        self.sizes_sum = 0.
        self.extra_lists = {}
        self.indices_list = []

    @property
    def loss(self):
        return 0 if self.sizes_sum == 0 else self.losses_sum / self.sizes_sum

    @property
    def metrics(self):
        if self.sizes_sum == 0:
            return dict(zip(self.metrics_names, np.zeros(len(self.metrics_names))))
        out = dict(zip(self.metrics_names, self.metrics_sum / self.sizes_sum))
        for i in range(self.nummodalities):
            names = ['%s_modal_%d' % (x, i) for x in self.metrics_names]
            out.update(dict(zip(names, self.metrics_permodal_sum[i] / self.sizes_sum)))
        return out

    @property
    def indices(self):
        if self.sizes_sum == 0 or self.indices_list[0] is None:
            return []
        return np.concatenate(self.indices_list, axis=0)

    def all_reduce(self, device=None):
        """Data parallel: make the epoch sums those of the GLOBAL batches so that every rank logs (and
        schedules the learning rate on) the same numbers.  One small collective per epoch and phase."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        flat = np.concatenate([[self.losses_sum, self.sizes_sum], self.metrics_sum, self.metrics_permodal_sum.ravel()])
        t = torch.from_numpy(flat).to(device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        flat = t.cpu().numpy()
        k = len(self.metrics_sum)
        self.losses_sum, self.sizes_sum = float(flat[0]), float(flat[1])
        self.metrics_sum = flat[2:2 + k].copy()
        self.metrics_permodal_sum = flat[2 + k:].reshape(self.metrics_permodal_sum.shape).copy()
        if self.indices_list and self.indices_list[0] is not None:
            # the utilization reader selects rows by `train_indices`: it needs every rank's
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, np.concatenate(self.indices_list, axis=0))
            self.indices_list = [np.concatenate(parts, axis=0)]

    def __iter__(self):
        if self.steps_per_epoch is not None:
            it = zip(range(1, self.steps_per_epoch + 1), _cycle(self.generator))
        else:
            it = zip(itertools.count(1), self.generator)
        for batch_ind, data in it:
            t0 = timeit.default_timer()
            self.callback.on_batch_begin(batch_ind, {})
            self.callback.on_forward_begin(batch_ind, data)
            step = {'number': batch_ind, 'indices': data[0]}
            yield step, data[1:]
            self.losses_sum += step['loss'] * step['size']
            self.metrics_sum += step['metrics'] * step['size']
            self.metrics_permodal_sum += step['viewwises_metrics'] * step['size']
            self.sizes_sum += step['size']
            idx = step['indices']
            self.indices_list.append(idx.cpu().numpy() if torch.is_tensor(idx) else idx)
            logs = dict(zip(self.metrics_names, step['metrics']))
            for i in range(self.nummodalities):
                names = ['%s_modal_%d' % (x, i) for x in self.metrics_names]
                logs.update(dict(zip(names, step['viewwises_metrics'][i])))
            for key, value in step.items():
                if key not in self.default_fields:
                    self.extra_lists.setdefault(key, []).append(value)
            self.callback.on_batch_end(batch_ind, {'batch': batch_ind, 'size': step['size'],
                                                   'time': timeit.default_timer() - t0, 'batch_begin_time': t0,
                                                   'loss': step['loss'], **logs})


class _Silent(Callback):
    pass


class DevicePrefetcher:
    """Input pipeline for the step engine: wraps a loader of `(idx, data, label)` batches and
    keeps ONE batch ahead on the device -- the host->device copy of batch i+1 (from pinned memory, on a side
    stream) overlaps the compute of batch i.  Tuple layout and order are unchanged (dataset.py:116-128)."""

    def __init__(self, loader, device, pin=True):
        self.loader, self.device, self.pin = loader, torch.device(device), pin
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        idx, data, label = batch
        conv = lambda t: torch.from_numpy(t) if isinstance(t, np.ndarray) else t
        data, label = conv(data), conv(label)
        if self.pin and not data.is_cuda and not data.is_pinned():
            data, label = data.pin_memory(), label.pin_memory()
        with torch.cuda.stream(self.stream):
            d = data.to(self.device, non_blocking=True)
            l = label.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return idx, d, l, ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            idx, d, l, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            d.record_stream(cur)
            l.record_stream(cur)
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            yield idx, d, l


class Model_:
    """Train / eval loops; reference src/framework.py:125-345.

    `metrics` may be `[acc]` from this module (fast path: device-side counts, one read-back
    per step) or arbitrary callables `(pred, y) -> scalar` (generic path, one sync each, as in
    the reference).  `data_parallel` is an optional greedy_multimodal_learning_b200.dist
    `GradientAllReduce`; its reduction finishes before on_backward_end so the learning-speed
    statistic sees the all-reduced gradients on every rank.
    """

    def __init__(self, model, optimizer, loss_function, nummodalities, *, metrics=[], verbose=True,
                 data_parallel=None, hyper_optim=None, vg=None):
        self.model, self.optimizer, self.loss_function = model, optimizer, loss_function
        self.metrics = metrics
        self.metrics_names = [m.__name__ for m in metrics]
        self.device = None
        self.verbose = verbose
        self.nummodalities = nummodalities
        self.curation_mode = False
        self.caring_modality = None
        self.data_parallel = data_parallel
        # device-side counts recompute the fused prediction as (l0 + l1) / 2 (src/model.py:108): only valid for models
        # that say so (`fused_logits_are_view_mean`, set by MMTM_MVCNN); anything else takes the generic path
        self._fast_acc = (len(metrics) == 1 and metrics[0] is acc and
                          bool(getattr(model, "fused_logits_are_view_mean", False)))
        self.last_correct_counts = None
        self._stat_dev = None

    # -- device plumbing ------------------------------------------------------------------
    def to(self, device):
        self.device = device
        self.model.to(device)
        return self

    def _process_input(self, x, y):
        conv = lambda t: torch.from_numpy(t) if isinstance(t, np.ndarray) else t
        x, y = conv(x), conv(y)
        if self.device is not None:
            x = x.to(self.device, non_blocking=True)
            y = y.to(self.device, non_blocking=True)
        return x, y

    # -- one batch ------------------------------------------------------------------------
    def _forward_loss(self, x, y):
        x, y = self._process_input(x, y)
        self.minibatch_data = (x, y)
        pred_eval, pred_y, scales, squeezed = self.model(x, curation_mode=self.curation_mode,
                                                         caring_modality=self.caring_modality)
        loss = self.loss_function(pred_y, y)
        record = {}
        if getattr(self.model, "saving_mmtm_scales", False):
            record['mmtmscales_list'] = scales
        if getattr(self.model, "saving_mmtm_squeeze_array", False):
            record['squeezedmaps_array_list'] = squeezed
        return loss, pred_eval, pred_y, y, record

    def _launch_counts(self, pred_y, y):
        """Queue the accuracy-count kernel; returns the device int32[3] tensor."""
        lib = _lib.load()
        l0, l1 = pred_y[0].detach().contiguous(), pred_y[1].detach().contiguous()
        _lib.require_cuda(l0, l1, y)
        if l0.dtype != torch.float32 or y.dtype != torch.int64:
            raise NotImplementedError("accuracy counts: fp32 logits and int64 labels only")
        counts = torch.empty(3, dtype=torch.int32, device=l0.device)
        with torch.cuda.device(l0.device):
            _lib.check(lib.gml_accuracy_counts(l0.data_ptr(), l1.data_ptr(), y.contiguous().data_ptr(), l0.shape[0],
                                               l0.shape[1], counts.data_ptr(), _lib.current_stream(l0.device)),
                       "gml_accuracy_counts")
        return counts

    def _read_back(self, loss, counts, n):
        """ONE device->host transfer for loss + counts."""
        dev = loss.device
        if self._stat_dev is None or self._stat_dev.device != dev:
            self._stat_dev = torch.empty(4, dtype=torch.float64, device=dev)
            self._stat_host = torch.empty(4, dtype=torch.float64).pin_memory()
        self._stat_dev[0] = loss.detach().double()
        self._stat_dev[1:] = counts.double()
        self._stat_host.copy_(self._stat_dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        v = self._stat_host.tolist()
        # float(loss_tensor) in the reference is the fp32 value widened to double
        k = [int(round(c)) for c in v[1:]]
        metrics = np.array([percent_from_count(k[0], n)])
        view = np.array([[percent_from_count(k[1], n)], [percent_from_count(k[2], n)]])
        return v[0], metrics, view, k

    def _generic_metrics(self, pred_eval, pred_y, y):
        m = lambda p: np.array([float(metric(p, y)) for metric in self.metrics])
        return m(pred_eval), np.array([m(p) for p in pred_y])

    def _finish_step(self, step, loss, pred_eval, pred_y, y, record):
        n = step['size']
        with torch.no_grad():
            if self._fast_acc and self.nummodalities == 2 and loss.is_cuda:
                counts = self._launch_counts(pred_y, y)
                step['loss'], step['metrics'], step['viewwises_metrics'], self.last_correct_counts = \
                    self._read_back(loss, counts, n)
            else:
                step['metrics'], step['viewwises_metrics'] = self._generic_metrics(pred_eval, pred_y, y)
                step['loss'] = float(loss)
        step.update(record)

    @staticmethod
    def _batch_size(x, y):
        if torch.is_tensor(x) or isinstance(x, np.ndarray):
            return len(x)
        if torch.is_tensor(y) or isinstance(y, np.ndarray):
            return len(y)
        return 1

    # -- loops ----------------------------------------------------------------------------
    def _eval_generator(self, generator, phase, *, steps=None):
        if steps is None:
            steps = len(generator)
        it = StepIterator(generator, steps, CallbackList([_Silent()]), self.metrics_names, self.nummodalities)
        self.model.eval()
        with torch.no_grad():
            for step, (x, y) in it:
                step['size'] = self._batch_size(x, y)
                loss, pred_eval, pred_y, yd, record = self._forward_loss(x, y)
                self._finish_step(step, loss, pred_eval, pred_y, yd, record)
        if self.data_parallel is not None:
            it.all_reduce(self.device)
        info = {'%s_loss' % phase: it.loss, '%s_indices' % phase: it.indices,
                **{'%s_%s' % (phase, k): v for k, v in it.extra_lists.items()},
                **{'%s_%s' % (phase, k): v for k, v in it.metrics.items()}}
        return info

    def eval_loop(self, test_generator, *, test_steps=None, epochs=1, callbacks=[]):
        cbs = CallbackList(callbacks)
        cbs.set_model_pytoune(self)
        cbs.on_train_begin({})
        epoch = 0
        while epoch <= epochs:
            t0 = timeit.default_timer()
            cbs.on_epoch_begin(epoch, {})
            logs = self._eval_generator(test_generator, 'test', steps=test_steps)
            logs.update(epoch=epoch, time=timeit.default_timer() - t0, epoch_begin_time=t0)
            cbs.on_epoch_end(epoch, logs)
            epoch += 1

    def train_step(self, step, x, y, callback_list):
        """One optimisation step (reference src/framework.py:307-322)."""
        step['size'] = self._batch_size(x, y)
        self.optimizer.zero_grad()
        loss, pred_eval, pred_y, yd, record = self._forward_loss(x, y)
        loss.backward()
        if self.data_parallel is not None:
            self.data_parallel.finish()
        callback_list.on_backward_end(step['number'])
        self.optimizer.step()
        self._finish_step(step, loss, pred_eval, pred_y, yd, record)
        if math.isnan(step['loss']):
            self.stop_training = True

    def train_loop(self, train_generator, test_generator=None, valid_generator=None, *, epochs=1000,
                   steps_per_epoch=None, validation_steps=None, test_steps=None, callbacks=[]):
        cbs = CallbackList(callbacks)
        cbs.set_model_pytoune(self)
        cbs.set_params({'epochs': epochs, 'steps': steps_per_epoch})
        self.stop_training = False
        cbs.on_train_begin({})
        history = []
        for epoch in range(1, epochs + 1):
            cbs.on_epoch_begin(epoch, {})
            t0 = timeit.default_timer()
            it = StepIterator(train_generator, steps_per_epoch, cbs, self.metrics_names, self.nummodalities)
            self.model.train(True)
            with torch.enable_grad():
                for step, (x, y) in it:
                    self.train_step(step, x, y, cbs)
            if self.data_parallel is not None:
                it.all_reduce(self.device)
            logs = {'loss': it.loss, 'train_indices': it.indices,
                    **{'train_%s' % k: v for k, v in it.extra_lists.items()}, **it.metrics}
            val = self._eval_generator(valid_generator, 'val', steps=validation_steps) if valid_generator is not None else {}
            test = self._eval_generator(test_generator, 'test', steps=test_steps) if test_generator is not None else {}
            epoch_log = {'epoch': epoch, 'time': timeit.default_timer() - t0, 'epoch_begin_time': t0, **logs, **val,
                         **test}
            cbs.on_epoch_end(epoch, epoch_log)
            history.append(epoch_log)
            if self.stop_training:
                break
        cbs.on_train_end({})
        return history
