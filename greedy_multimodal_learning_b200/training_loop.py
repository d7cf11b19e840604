"""Epoch driver and history files: mirror of the reference's `training_loop` / `evalution_loop`
(src/training_loop.py:86-143,163-215) on top of this package's step engine.

File formats are the reference's, because its own tools read them back
(SURVEY 8f-2): `history.csv` (scalar columns only), `history.pickle` (the whole history dict,
including the per-batch lists of CPU squeeze tensors that `get_mmtm_outputs` walks,
src/balanced_mmtm.py:157-176), `model_best_val.pt` / `model_last_epoch.pt`
(`{'model':…, 'optimizer':…}`), and for evaluation `<save>/eval_history_batch/history.*`.
"""
from __future__ import annotations

import csv
import logging
import os
import pickle
from functools import partial

import numpy as np
import torch

from . import dist as gdist
from . import gin_lite
from .callbacks import LambdaCallback, ModelCheckpoint, save_weights
from .framework import Model_

logger = logging.getLogger(__name__)

_CSV_TYPES = (int, float, complex, np.integer, np.floating, str)


def _append_to_history(epoch, logs, H):
    for key, value in logs.items():
        H.setdefault(key, []).append(value)


def _save_history(epoch, logs, save_path, H, save_with_structure=False):
    logger.info("\t".join("%s=%s" % (k, v) for k, v in logs.items() if isinstance(v, _CSV_TYPES)))
    cols = {k: v for k, v in H.items() if isinstance(v[-1], _CSV_TYPES)}
    with open(os.path.join(save_path, "history.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(list(cols))
        for row in zip(*cols.values()):
            w.writerow(row)
    if save_with_structure:
        with open(os.path.join(save_path, "history.pickle"), "wb") as f:
            pickle.dump(H, f, pickle.HIGHEST_PROTOCOL)


def _remove(*paths):
    for p in paths:
        if os.path.exists(p):
            os.remove(p)


def _load_pretrained_model(model, path):
    """Non-strict reload of `checkpoint['model']` (src/training_loop.py:78-83)."""
    checkpoint = torch.load(path, map_location="cpu")
    state = model.state_dict()
    state.update(checkpoint['model'])
    model.load_state_dict(state, strict=False)


def _configure(callbacks, save_path, model, optimizer, config):
    for cb in callbacks:
        cb.set_save_path(save_path)
        cb.set_model(model, ignore=False)
        if optimizer is not None:
            cb.set_optimizer(optimizer)
        cb.set_config(config)


def _device(use_gpu, device_numbers):
    if not use_gpu:
        return None
    if not torch.cuda.is_available():
        raise RuntimeError("use_gpu=True but no CUDA device is visible (this package has no CPU path)")
    if gdist.world_size() > 1:  # one process per GPU: the launcher's LOCAL_RANK picks the device
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cuda:%d" % device_numbers[0])


@gin_lite.configurable
def training_loop(model, loss_function, metrics, optimizer, config, save_path, steps_per_epoch, train=None,
                  valid=None, test=None, test_steps=None, validation_steps=None, use_gpu=False, device_numbers=[0],
                  custom_callbacks=[], checkpoint_monitor="val_acc", n_epochs=100, verbose=True, nummodalities=2,
                  data_parallel=None):
    _remove(os.path.join(save_path, "history.pkl"), os.path.join(save_path, "history.csv"))
    H = {}
    callbacks = list(custom_callbacks)
    # the reference passes `custom_callbacks` (a list) as `save_with_structure`, so the pickle is
    # written exactly when custom callbacks exist (src/training_loop.py:108-109); kept
    callbacks.append(LambdaCallback(on_epoch_end=partial(_append_to_history, H=H)))
    if gdist.rank() == 0:  # data parallel: replicas are identical, rank 0 owns the files
        callbacks += [
            LambdaCallback(on_epoch_end=partial(_save_history, save_path=save_path, H=H,
                                                save_with_structure=bool(custom_callbacks))),
            ModelCheckpoint(monitor=checkpoint_monitor, save_best_only=True, mode='max',
                            filepath=os.path.join(save_path, "model_best_val.pt")),
            LambdaCallback(on_epoch_end=lambda epoch, logs: save_weights(
                model, optimizer, os.path.join(save_path, "model_last_epoch.pt"))),
        ]
    _configure(callbacks, save_path, model, optimizer, config)
    engine = Model_(model=model, optimizer=optimizer, loss_function=loss_function, metrics=metrics, verbose=verbose,
                    nummodalities=nummodalities, data_parallel=data_parallel)
    for cb in callbacks:
        cb.set_model_pytoune(engine)
    dev = _device(use_gpu, device_numbers)
    if dev is not None:
        engine.to(dev)
    # `epochs=n_epochs - 1` is the reference's own off-by-one (src/training_loop.py:140)
    engine.train_loop(train, valid_generator=valid, test_generator=test, test_steps=test_steps,
                      validation_steps=validation_steps, steps_per_epoch=steps_per_epoch, epochs=n_epochs - 1,
                      callbacks=callbacks)
    return H


@gin_lite.configurable
def evalution_loop(model, loss_function, metrics, config, save_path, test=None, test_steps=None, use_gpu=False,
                   device_numbers=[0], custom_callbacks=[], pretrained_weights_path=None, save_with_structure=False,
                   nummodalities=2):
    """(sic) -- the reference's spelling is the gin name the config files bind."""
    if pretrained_weights_path is not None:
        _load_pretrained_model(model, pretrained_weights_path)
    _remove(os.path.join(save_path, "eval_history.pkl"), os.path.join(save_path, "eval_history.csv"))
    history_batch = os.path.join(save_path, 'eval_history_batch')
    os.makedirs(history_batch, exist_ok=True)
    H = {}
    callbacks = list(custom_callbacks) + [
        LambdaCallback(on_epoch_end=partial(_append_to_history, H=H)),
        LambdaCallback(on_epoch_end=partial(_save_history, save_path=history_batch, H=H,
                                            save_with_structure=save_with_structure)),
    ]
    _configure(callbacks, save_path, model, None, config)
    engine = Model_(model=model, optimizer=None, loss_function=loss_function, metrics=metrics,
                    nummodalities=nummodalities)
    for cb in callbacks:
        cb.set_model_pytoune(engine)
    dev = _device(use_gpu, device_numbers)
    if dev is not None:
        engine.to(dev)
    engine.eval_loop(test, epochs=0, test_steps=test_steps, callbacks=callbacks)
    return H
