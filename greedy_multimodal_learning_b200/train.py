"""`python -m greedy_multimodal_learning_b200.train <save_path> <config.gin> [-b bindings]` --
the reference's `train.py` entry point (train.py:43-75) on the CUDA path.  Launched under
`torch.distributed.run` it trains data-parallel (batch sharded over ranks, see dist.py)."""
from __future__ import annotations

import torch

from . import callbacks as avail_callbacks
from . import dataset, dist, gin_lite
from .framework import DevicePrefetcher, acc, blend_loss
from .model import MMTM_MVCNN
from .training_loop import training_loop
from .utils import gin_wrap


@gin_lite.configurable
def train(save_path, wd, lr, momentum, batch_size, callbacks=[]):
    multi = dist.init_from_env()  # torchrun environment -> one process per GPU
    torch.backends.cudnn.benchmark = True  # fixed input shapes: let cuDNN time its convolution algorithms once
    dev = torch.device("cuda:%d" % torch.cuda.current_device()) if torch.cuda.is_available() else None
    model = MMTM_MVCNN()
    train_loader, valid, test = dataset.get_mvdcndata(batch_size=batch_size)
    make_opt = lambda params: torch.optim.SGD(params, lr=lr, weight_decay=wd, momentum=momentum)
    constructed = [avail_callbacks.__dict__[name]() for name in callbacks if name in avail_callbacks.__dict__]
    data_parallel = None
    if multi:
        # every rank draws the same global batch (same seed) and keeps its contiguous slice
        model.to(dev)
        model, optimizer, data_parallel = dist.setup_model(model, make_opt)
        train_loader, valid, test = (dist.ShardedBatches(l) for l in (train_loader, valid, test))
    else:
        optimizer = make_opt(model.parameters())
    n_train, n_valid, n_test = len(train_loader), len(valid), len(test)
    if dev is not None:
        train_loader, valid, test = (DevicePrefetcher(l, dev) for l in (train_loader, valid, test))
    return training_loop(model=model, optimizer=optimizer, loss_function=blend_loss, metrics=[acc],
                         train=train_loader, valid=valid, test=test, steps_per_epoch=n_train,
                         validation_steps=n_valid, test_steps=n_test, save_path=save_path,
                         config=gin_lite.config_dict(), custom_callbacks=constructed, data_parallel=data_parallel)


if __name__ == "__main__":
    gin_wrap(train)
