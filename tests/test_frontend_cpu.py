"""Front end around the path: gin-compatible parser, history files, auxiliary
callbacks, dataset tuple layout.  Host logic only -- nothing here computes on a GPU."""
import csv
import glob
import json
import os
import pickle

import numpy as np
import pytest
import torch

from greedy_multimodal_learning_b200 import callbacks as cbs
from greedy_multimodal_learning_b200 import dataset, gin_lite, training_loop

GUIDED = """
# Model
MMTM_MVCNN.pretraining=False
MMTM_MVCNN.num_views=2

# Train configuration
train.batch_size=8
train.lr=0.1     # trailing comment
train.callbacks=['CompletedStopping', 'ReduceLROnPlateau_PyTorch', 'Bias_Mitigation_Strong']
Bias_Mitigation_Strong.epsilon=0.01
Bias_Mitigation_Strong.MMTMnames = ['visual', 'skeleton']
eval_.pretrained_weights_path='/some/path#with_hash/model_best_val.pt'
get_mvdcndata.specific_views=[0,
                              6]
"""


@pytest.fixture(autouse=True)
def _clean_gin():
    gin_lite.clear_config()
    yield
    gin_lite.clear_config()


def test_parser_handles_the_reference_syntax():
    gin_lite.parse_config(GUIDED)
    c = gin_lite.config_dict()
    assert c["train.lr"] == 0.1 and c["train.batch_size"] == 8 and c["MMTM_MVCNN.pretraining"] is False
    assert c["train.callbacks"][-1] == "Bias_Mitigation_Strong"
    assert c["Bias_Mitigation_Strong.MMTMnames"] == ["visual", "skeleton"]
    assert c["eval_.pretrained_weights_path"].endswith("#with_hash/model_best_val.pt")  # '#' inside a string
    assert c["get_mvdcndata.specific_views"] == [0, 6]
    gin_lite.parse_config_files_and_bindings([], "train.lr=0.5\ntrain.wd=1e-4")  # bindings override files
    assert gin_lite.query_parameter("train.lr") == 0.5 and gin_lite.query_parameter("train.wd") == 1e-4


@pytest.mark.parametrize("bad", ["train.lr = @foo()", "train.lr = %MACRO", "scope/train.lr = 1", "import a.b",
                                 "include 'x.gin'"])
def test_unsupported_gin_features_fail_loudly(bad):
    with pytest.raises(NotImplementedError):
        gin_lite.parse_config(bad)


def test_malformed_statements_raise():
    for bad in ("just words", "lr = 3", "a.b = [1, 2"):
        with pytest.raises((ValueError, SyntaxError)):
            gin_lite.parse_config(bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/configs"), reason="reference tree not mounted")
def test_every_reference_config_parses_and_binds_known_parameters():
    import inspect
    from greedy_multimodal_learning_b200 import eval as eval_mod, model, train as train_mod
    targets = {"MMTM_MVCNN": model.MMTM_MVCNN.__init__.__wrapped__, "train": train_mod.train.__wrapped__,
               "eval_": eval_mod.eval_.__wrapped__, "training_loop": training_loop.training_loop.__wrapped__,
               "evalution_loop": training_loop.evalution_loop.__wrapped__,
               "get_mvdcndata": dataset.get_mvdcndata.__wrapped__,
               "CompletedStopping": cbs.CompletedStopping.__init__.__wrapped__,
               "ReduceLROnPlateau_PyTorch": cbs.ReduceLROnPlateau_PyTorch.__init__.__wrapped__,
               "Bias_Mitigation_Strong": cbs.Bias_Mitigation_Strong.__init__.__wrapped__,
               "ProgressionCallback": cbs.ProgressionCallback.__init__.__wrapped__}
    files = sorted(glob.glob("/root/reference/configs/*.gin"))
    assert len(files) == 5
    for f in files:
        gin_lite.clear_config()
        gin_lite.parse_config_files_and_bindings([f])
        for key in gin_lite.config_dict():
            name, param = key.rsplit(".", 1)
            assert name in targets, "%s: unknown configurable %s" % (f, name)
            assert param in inspect.signature(targets[name]).parameters, "%s: %s" % (f, key)


def test_configurable_fills_only_what_the_caller_left_out():
    @gin_lite.configurable
    def f(a, b=2, c=3):
        return a, b, c

    @gin_lite.configurable("Other")
    class K:
        def __init__(self, x, y=0):
            self.x, self.y = x, y

    gin_lite.parse_config("f.b=20\nf.c=30\nOther.x='gx'\nOther.y=5")
    assert f(1) == (1, 20, 30) and f(1, 7) == (1, 7, 30) and f(1, c=9) == (1, 20, 9)
    assert (K().x, K().y, K("mine").x, K(y=1).y) == ("gx", 5, "mine", 1)
    gin_lite.bind_parameter("f.nope", 1)
    with pytest.raises(TypeError):
        f(1)


def test_guided_controller_takes_its_arguments_from_the_config():
    gin_lite.parse_config("Bias_Mitigation_Strong.epsilon=0.01\nBias_Mitigation_Strong.curation_windowsize=5\n"
                          "Bias_Mitigation_Strong.starting_epoch=2\n"
                          "Bias_Mitigation_Strong.branchnames=['net_view_0', 'net_view_1']")
    cb = cbs.__dict__["Bias_Mitigation_Strong"]()  # how train.py constructs callbacks (train.py:54-58)
    assert (cb.epsilon, cb.curation_windowsize, cb.starting_epoch) == (0.01, 5, 2)
    assert cb.MMTMnames == ['visual', 'skeleton']


def test_history_files_keep_the_reference_layout(tmp_path):
    H = {}
    sq = [[[torch.ones(2, 4), torch.zeros(2, 4)]] * 3]  # per batch: 3 blocks x 2 views
    for epoch, a in enumerate((50.0, np.float32(75.0)), 1):
        logs = {"epoch": epoch, "acc": a, "loss": 1.0 / epoch, "train_indices": np.arange(4),
                "test_squeezedmaps_array_list": sq, "caring_modality": None if epoch == 1 else 0}
        training_loop._append_to_history(epoch, logs, H)
        training_loop._save_history(epoch, logs, str(tmp_path), H, save_with_structure=True)
    rows = list(csv.DictReader(open(tmp_path / "history.csv")))
    assert [r["epoch"] for r in rows] == ["1", "2"] and float(rows[1]["acc"]) == 75.0
    assert "train_indices" not in rows[0] and "test_squeezedmaps_array_list" not in rows[0]
    back = pickle.load(open(tmp_path / "history.pickle", "rb"))
    assert np.array_equal(back["train_indices"][0], np.arange(4))
    assert torch.equal(back["test_squeezedmaps_array_list"][0][0][2][0], torch.ones(2, 4))


def test_history_pickle_round_trips_through_get_rescale_weights(tmp_path):
    """Writer (this package) -> reader (`get_rescale_weights`, same format as the reference's reader)."""
    from greedy_multimodal_learning_b200 import get_rescale_weights
    rs = np.random.RandomState(0)
    n, dims = 10, (8, 16, 32)
    full = [[rs.standard_normal((n, c)).astype(np.float32) for _ in range(2)] for c in dims]
    perm = rs.permutation(n)
    batches, idx = [], []
    for lo in range(0, n, 4):
        sel = perm[lo:lo + 4]
        idx.append(sel)
        batches.append([[torch.from_numpy(full[b][v][sel]) for v in range(2)] for b in range(3)])
    ev, tr = tmp_path / "eval_history_batch", tmp_path
    ev.mkdir()
    He, Ht = {}, {}
    training_loop._append_to_history(0, {"test_indices": np.concatenate(idx), "test_squeezedmaps_array_list": batches,
                                         "test_loss": 1.0}, He)
    training_loop._save_history(0, {}, str(ev), He, save_with_structure=True)
    train_idx = np.array([1, 3, 4, 8])
    training_loop._append_to_history(1, {"train_indices": train_idx, "loss": 1.0}, Ht)
    training_loop._save_history(1, {}, str(tr), Ht, save_with_structure=True)
    out = get_rescale_weights(str(ev), str(tr))
    assert out[0] is None and len(out) == 4
    for b in range(3):
        for v in range(2):
            np.testing.assert_allclose(out[b + 1][v], full[b][v][train_idx].mean(0), rtol=1e-6)


class _Engine:
    stop_training = False
    metrics_names = ["acc"]


def test_completed_stopping_counts_perfect_epochs_without_reset():
    cb, eng = cbs.CompletedStopping(monitor="acc", patience=2, verbose=False), _Engine()
    cb.set_model_pytoune(eng)
    cb.on_train_begin({})
    for epoch, a in enumerate((100, 99.0, 100.0), 1):
        cb.on_epoch_end(epoch, {"acc": a})
        assert eng.stop_training == (epoch == 3)
    assert cb.stopped_epoch == 3


def test_model_checkpoint_saves_only_improvements(tmp_path):
    m = torch.nn.Linear(2, 2)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    path = tmp_path / "model_best_val.pt"
    cb = cbs.ModelCheckpoint(str(path), monitor="val_acc", save_best_only=True, mode="max")
    cb.set_model(m, ignore=False)
    cb.set_optimizer(opt)
    cb.on_epoch_end(1, {"val_acc": 10.0})
    first = os.path.getmtime(path)
    with torch.no_grad():
        m.weight.add_(1.0)
    cb.on_epoch_end(2, {"val_acc": 5.0})
    kept = torch.load(path)
    assert set(kept) == {"model", "optimizer"} and not torch.equal(kept["model"]["weight"], m.weight)
    cb.on_epoch_end(3, {"val_acc": 11.0})
    assert torch.equal(torch.load(path)["model"]["weight"], m.weight) and os.path.getmtime(path) >= first
    cb.on_epoch_end(4, {})  # monitor missing -> skipped, no exception


def test_reduce_lr_on_plateau_and_lambda_callback():
    m = torch.nn.Linear(2, 2)
    opt = torch.optim.SGD(m.parameters(), lr=1.0)
    cb = cbs.ReduceLROnPlateau_PyTorch("loss", factor=0.5, patience=0)
    cb.set_optimizer(opt)
    cb.on_train_begin({})
    for epoch in range(3):
        cb.on_epoch_end(epoch, {"loss": 1.0})
    assert opt.param_groups[0]["lr"] == 0.25
    seen = []
    lc = cbs.LambdaCallback(on_epoch_end=lambda e, logs: seen.append(e))
    lc.on_epoch_begin(1, {})
    lc.on_epoch_end(1, {})
    assert seen == [1]
    with pytest.raises(TypeError):
        cbs.LambdaCallback(on_nothing=print)


def test_synthetic_loaders_keep_the_reference_tuple_and_split():
    tr, va, te = dataset.get_mvdcndata(batch_size=4, synthetic_samples=(20, 6), image_size=16,
                                       specific_views=[0, 6], num_views=12, use_cuda=False)
    a = next(iter(tr))  # the shuffle draws from the global RNG that get_mvdcndata just seeded
    assert (len(tr.dataset), len(va.dataset), len(te.dataset)) == (16, 4, 6)  # valid_size 0.2
    idx, data, label = next(iter(te))
    assert data.shape == (4, 2, 3, 16, 16) and data.dtype == torch.float32 and label.dtype == torch.int64
    assert idx.tolist() == [0, 1, 2, 3]
    tr2, va2, _ = dataset.get_mvdcndata(batch_size=4, synthetic_samples=(20, 6), image_size=16,
                                        specific_views=[0, 6], use_cuda=False)
    assert va.dataset.indices == va2.dataset.indices  # split depends on random_seed_for_validation only
    b = next(iter(tr2))                               # same seed -> same shuffled first batch
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    seen = sorted(int(i) for batch in tr for i in batch[0])
    assert seen == sorted(tr.dataset.indices) and not set(seen) & set(va.dataset.indices)


def test_modelnet_directory_layout_is_read(tmp_path):
    classes = ["chair", "sofa"]
    meta = {"classnames": classes, "train": [], "test": []}
    rs = np.random.RandomState(0)
    for split, n in (("train", 5), ("test", 2)):
        (tmp_path / split).mkdir()
        for i in range(n):
            name = "%s_%04d" % (split, i)
            meta[split].append({"classname": classes[i % 2], "model": name})
            torch.save(rs.randint(0, 255, (12, 8, 8, 3)).astype(np.uint8), tmp_path / split / (name + ".npy"))
    json.dump(meta, open(tmp_path / "metadata.json", "w"))
    tr, va, te = dataset.get_mvdcndata(root_dir=str(tmp_path), batch_size=2, specific_views=[0, 6], valid_size=0.2,
                                       use_cuda=False)
    idx, data, label = next(iter(te))
    assert data.shape == (2, 2, 3, 8, 8) and label.tolist() == [0, 1] and idx.tolist() == [0, 1]
    raw = torch.load(tmp_path / "test" / "test_0000.npy", weights_only=False)[6].astype(np.float32) / 255.0
    want = (torch.from_numpy(raw).permute(2, 0, 1) - dataset._MEAN) / dataset._STD
    torch.testing.assert_close(data[0, 1], want)
    assert len(tr.dataset) == 4 and len(va.dataset) == 1


def test_train_record_eval_file_plumbing_with_the_oracle_block(tmp_path):
    """Same flow as tests/test_frontend_gpu.py with the oracle block plugged in (CPU, test-only)."""
    from tests.frontend_flow import run_flow
    from oracle.mmtm_module import OracleMMTM
    run_flow(tmp_path, "cpu", OracleMMTM)
