"""Shared body of the front-end flow tests: the reference's three command lines (README.md:14-23) with its
config bindings on synthetic data.  `device='cuda:0'` runs the product; `device='cpu'` plugs the ORACLE block
into the same host code (tests only) so the file plumbing is covered without a GPU."""
import csv
import os
import pickle

import numpy as np
import torch

DATA = """
get_mvdcndata.make_npy_files=False
get_mvdcndata.num_views=2
get_mvdcndata.num_workers=0
get_mvdcndata.specific_views=[0, 6]
get_mvdcndata.synthetic_samples=(24, 8)
get_mvdcndata.image_size=64
"""

TRAIN = """
MMTM_MVCNN.pretraining=False
MMTM_MVCNN.num_views=2
train.batch_size=4
train.lr=0.01
train.wd=0.0
train.momentum=0
train.callbacks=%(callbacks)s
ReduceLROnPlateau_PyTorch.metric='loss'
CompletedStopping.patience=5
CompletedStopping.monitor='acc'
Bias_Mitigation_Strong.epsilon=0.01
Bias_Mitigation_Strong.curation_windowsize=5
Bias_Mitigation_Strong.starting_epoch=2
Bias_Mitigation_Strong.branchnames=['net_view_0', 'net_view_1']
Bias_Mitigation_Strong.MMTMnames = ['visual', 'skeleton']
training_loop.nummodalities=2
training_loop.n_epochs=4
training_loop.use_gpu=%(use_gpu)s
training_loop.device_numbers=[0]
training_loop.checkpoint_monitor='val_acc'
""" + DATA

RECORD = """
MMTM_MVCNN.pretraining=False
MMTM_MVCNN.num_views=2
MMTM_MVCNN.saving_mmtm_squeeze_array=True
eval_.target_data_split='train'
eval_.batch_size=8
eval_.pretrained_weights_path='%(train)s/model_best_val.pt'
evalution_loop.use_gpu=%(use_gpu)s
evalution_loop.device_numbers=[0]
evalution_loop.save_with_structure=True
get_mvdcndata.valid_size=0
""" + DATA

EVAL = """
MMTM_MVCNN.pretraining=False
MMTM_MVCNN.num_views=2
MMTM_MVCNN.mmtm_off=True
MMTM_MVCNN.mmtm_rescale_eval_file_path='%(record)s/eval_history_batch'
MMTM_MVCNN.mmtm_rescale_training_file_path='%(train)s'
MMTM_MVCNN.device='%(device)s'
eval_.target_data_split='test'
eval_.batch_size=8
eval_.pretrained_weights_path='%(train)s/model_best_val.pt'
evalution_loop.use_gpu=%(use_gpu)s
evalution_loop.device_numbers=[0]
evalution_loop.save_with_structure=False
""" + DATA


def _run(fn, save, cfg_text, tmp_path, name, mmtm_cls):
    from greedy_multimodal_learning_b200 import gin_lite
    from greedy_multimodal_learning_b200.utils import gin_wrap
    cfg = tmp_path / (name + ".gin")
    cfg.write_text(cfg_text)
    gin_lite.clear_config()
    if mmtm_cls is not None:
        gin_lite.bind_parameter('MMTM_MVCNN.mmtm_cls', mmtm_cls)
    try:
        return gin_wrap(fn, [str(save), str(cfg)])
    finally:
        gin_lite.clear_config()


def run_flow(tmp_path, device, mmtm_cls=None):
    from greedy_multimodal_learning_b200 import get_rescale_weights
    from greedy_multimodal_learning_b200.eval import eval_
    from greedy_multimodal_learning_b200.train import train
    paths = {k: str(tmp_path / k) for k in ("train", "record", "eval")}
    guided = device != "cpu"  # the learning-speed statistic has no CPU path
    paths.update(device=device, use_gpu=str(device != "cpu"),
                 callbacks=str(['CompletedStopping', 'ReduceLROnPlateau_PyTorch'] +
                               (['Bias_Mitigation_Strong'] if guided else [])))

    # 1. training_guided.gin: 3 epochs (the reference's n_epochs - 1) of 5 steps
    _run(train, paths["train"], TRAIN % paths, tmp_path, "training_guided", mmtm_cls)
    for f in ("history.csv", "history.pickle", "model_best_val.pt", "model_last_epoch.pt", "stdout.txt", "stderr.txt"):
        assert os.path.exists(os.path.join(paths["train"], f)), f
    rows = list(csv.DictReader(open(os.path.join(paths["train"], "history.csv"))))
    assert [int(r["epoch"]) for r in rows] == [1, 2, 3]
    for key in ("loss", "acc", "acc_modal_0", "acc_modal_1", "val_loss", "val_acc", "test_acc", "time"):
        assert key in rows[0] and np.isfinite(float(rows[-1][key])), key
    hist = pickle.load(open(os.path.join(paths["train"], "history.pickle"), "rb"))
    assert sorted(hist["train_indices"][0].tolist()) == sorted(set(hist["train_indices"][0].tolist()))
    assert len(hist["train_indices"][0]) == 20 and len(hist["val_indices"][0]) == 4  # 24 samples, valid_size 0.2
    ckpt = torch.load(os.path.join(paths["train"], "model_best_val.pt"), map_location="cpu")
    assert "mmtm4.fc_squeeze.weight" in ckpt["model"] and "net_view_1.fc.bias" in ckpt["model"]

    # 2. recording.gin: squeezes of every training sample into eval_history_batch/history.pickle
    _run(eval_, paths["record"], RECORD % paths, tmp_path, "recording", mmtm_cls)
    rec = pickle.load(open(os.path.join(paths["record"], "eval_history_batch", "history.pickle"), "rb"))
    batches = rec["test_squeezedmaps_array_list"][0]
    assert len(batches) == 3 and len(batches[0]) == 3 and len(batches[0][0]) == 2
    assert [tuple(batches[0][b][0].shape) for b in range(3)] == [(8, 128), (8, 256), (8, 512)]
    assert not batches[0][0][0].is_cuda and sorted(rec["test_indices"][0].tolist()) == list(range(24))

    # 3. the reader turns the two pickles into per-block dataset means over the training indices
    means = get_rescale_weights(os.path.join(paths["record"], "eval_history_batch"), paths["train"])
    order = np.argsort(rec["test_indices"][0])
    sel = hist["train_indices"][0]
    for b in range(3):
        for v in range(2):
            full = np.concatenate([bt[b][v].numpy() for bt in batches])[order]
            np.testing.assert_allclose(means[b + 1][v], full[sel].mean(0), rtol=1e-6, atol=1e-7)

    # 4. eval.gin: cross-modal flow off, fed by those means
    _run(eval_, paths["eval"], EVAL % paths, tmp_path, "eval", mmtm_cls)
    rows = list(csv.DictReader(open(os.path.join(paths["eval"], "eval_history_batch", "history.csv"))))
    assert len(rows) == 1 and 0.0 <= float(rows[0]["test_acc"]) <= 100.0 and np.isfinite(float(rows[0]["test_loss"]))
    assert not os.path.exists(os.path.join(paths["eval"], "eval_history_batch", "history.pickle"))
