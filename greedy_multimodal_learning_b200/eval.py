"""`python -m greedy_multimodal_learning_b200.eval <save_path> <config.gin> [-b bindings]` -- the
reference's `eval.py` entry point (eval.py:24-60): one pass over a data split with a reloaded
checkpoint; `configs/recording.gin` records the squeeze arrays, `configs/eval.gin` evaluates with
the cross-modal flow replaced by the recorded dataset means (README.md:20-23)."""
from __future__ import annotations

import torch

from . import callbacks as avail_callbacks
from . import dataset, gin_lite
from .framework import acc, blend_loss
from .model import MMTM_MVCNN
from .training_loop import evalution_loop
from .utils import gin_wrap


@gin_lite.configurable("eval_")
def eval_(save_path, target_data_split, pretrained_weights_path, batch_size=128, callbacks=[]):
    model = MMTM_MVCNN()
    train, val, testing = dataset.get_mvdcndata(batch_size=batch_size)
    try:
        target_data = {'test': testing, 'train': train, 'val': val}[target_data_split]
    except KeyError:
        raise NotImplementedError(target_data_split)
    constructed = [avail_callbacks.__dict__[name]() for name in callbacks if name in avail_callbacks.__dict__]
    return evalution_loop(model=model, loss_function=blend_loss, metrics=[acc], config=gin_lite.config_dict(),
                          save_path=save_path, test=target_data, test_steps=len(target_data),
                          custom_callbacks=constructed, pretrained_weights_path=pretrained_weights_path)


if __name__ == "__main__":
    gin_wrap(eval_)
