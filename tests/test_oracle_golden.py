"""Pin the oracle restatement (oracle/*.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import mmtm_oracle as mo
from oracle import stats_oracle as so
from tests.helpers import assert_close, rel_err
from tests.golden import make_golden_cases as cases

G = os.path.join(os.path.dirname(__file__), "golden")


def _avg_for(seed, c):
    rs = np.random.RandomState(seed + 1000)
    return [torch.from_numpy((0.1 * rs.standard_normal(c)).astype(np.float32)) for _ in range(2)]


def _run_oracle(n, c, h, w, seed, mode, warm_n):
    x = mo.synth_inputs(seed, n, c, h, w)
    warm = mo.synth_inputs(seed + 500, warm_n, c, h, w)
    p = mo.synth_params(seed, c, c)
    st = mo.MMTMState.zeros(c)
    with torch.no_grad():
        mo.forward(warm["A"], warm["B"], p, st, mo.MODE_NORMAL)
    r = mo.forward_backward(x["A"], x["B"], p, st, x["gA"], x["gB"], mode, _avg_for(seed, c))
    r["run_v"], r["run_s"], r["step"] = st.run_v, st.run_s, st.step
    r64 = mo.forward_backward_f64(x["A"], x["B"], *[t.numpy() for t in p.tensors()], x["gA"], x["gB"],
                                  np.zeros(c), np.zeros(c), 0, mode, [a.numpy() for a in _avg_for(seed, c)])
    return r, r64


KEYS = ["A_out", "B_out", "dA", "dB", "gA", "gB", "dWsq", "dbsq", "dWv", "dbv", "dWs", "dbs", "run_v", "run_s"]


@pytest.mark.parametrize("case", cases.SMALL_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_small_shapes_full_tensors(case, mode):
    name, n, c, h, w, seed = case
    gold = np.load(os.path.join(G, "mmtm_small.npz"))
    r, r64 = _run_oracle(n, c, h, w, seed, mode, n + 1)
    for k in KEYS:
        assert_close(r[k], gold["%s/m%d/%s" % (name, mode, k)], 1e-5, "%s m%d %s" % (name, mode, k))
    assert int(gold["%s/m%d/step" % (name, mode)]) == r["step"] == 2
    if mode != 3:
        assert_close(r["sA"], gold["%s/m%d/sA" % (name, mode)], 1e-5, "sA")
        assert_close(r["sB"], gold["%s/m%d/sB" % (name, mode)], 1e-5, "sB")
    # which excitation FCs receive a gradient at all
    assert bool(gold["%s/m%d/wv_has_grad" % (name, mode)]) == r["has_grad"]["w_v"] == (mode != 1)
    assert bool(gold["%s/m%d/ws_has_grad" % (name, mode)]) == r["has_grad"]["w_s"] == (mode != 2)
    # the closed-form float64 backward agrees with the reference's autograd (prewarm does not
    # change gradients except through the substituted running mean, so only modes 0 and 3)
    if mode in (0, 3):
        for k in ["A_out", "B_out", "dA", "dB", "dWsq", "dbsq", "dWv", "dbv", "dWs", "dbs", "gA", "gB"]:
            assert rel_err(r64[k], gold["%s/m%d/%s" % (name, mode, k)]) < 2e-6, k


@pytest.mark.parametrize("case", cases.CONFIG_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_config_shapes_summaries(case, mode):
    name, n, c, h, seed = case
    gold = np.load(os.path.join(G, "mmtm_config.npz"))
    r, _ = _run_oracle(n, c, h, h, seed, mode, 3)
    for k in KEYS:
        key = "%s/m%d/%s" % (name, mode, k)
        v = r[k].detach().numpy() if isinstance(r[k], torch.Tensor) else np.asarray(r[k])
        if key in gold.files:
            assert_close(v, gold[key], 1e-5, key)
        else:
            flat = v.reshape(-1)
            assert_close(flat[::cases.SUB], gold[key + ".sub"], 1e-5, key + ".sub")
            scale = float(gold[key + ".abssum"])
            assert abs(flat.astype(np.float64).sum() - float(gold[key + ".sum"])) <= 1e-6 * max(scale, 1e-30)
            assert abs((flat.astype(np.float64) ** 2).sum() - float(gold[key + ".sqsum"])) <= 1e-5 * float(
                gold[key + ".sqsum"]) + 1e-30


def test_state_sequence():
    gold = np.load(os.path.join(G, "mmtm_sequence.npz"))
    c, h, seed = 16, 6, 31
    p = mo.synth_params(seed, c, c)
    st = mo.MMTMState.zeros(c)
    for i, (mode, n, grad) in enumerate(cases.SEQUENCE):
        x = mo.synth_inputs(seed + 10 * i, n, c, h)
        with torch.set_grad_enabled(grad):
            a_out, b_out, _ = mo.forward(x["A"], x["B"], p, st, mode)
        assert_close(a_out, gold["%d/A_out" % i], 1e-5, "A_out@%d" % i)
        assert_close(b_out, gold["%d/B_out" % i], 1e-5, "B_out@%d" % i)
        assert_close(st.run_v, gold["%d/run_v" % i], 1e-6, "run_v@%d" % i)
        assert_close(st.run_s, gold["%d/run_s" % i], 1e-6, "run_s@%d" % i)
        # reference quirk: both running means are fed by the visual gate -> identical
        assert torch.equal(st.run_v, st.run_s)
        assert st.step == int(gold["%d/step" % i]) == i + 1


def test_rescale_weights():
    gold = np.load(os.path.join(G, "rescale.npz"))
    ev, tr = cases.synth_history()
    for validation in (False, True):
        w = so.mean_squeezes_from_history(ev, tr, validation=validation)
        assert w[0] is None and len(w) == 4
        for pos in (1, 2, 3):
            for v in (0, 1):
                assert_close(w[pos][v], gold["val%d/pos%d/view%d" % (validation, pos, v)], 1e-6, "rescale")


def test_acc_and_loss():
    for case in json.load(open(os.path.join(G, "acc.json"))):
        lt, l2 = torch.tensor(case["logits"]), torch.tensor(case["logits2"])
        y = torch.tensor(case["y"])
        assert float(so.acc(lt, y)) == case["acc"]              # bit-exact: integer counts / n * 100
        assert float(so.acc([lt, l2], y)) == case["acc_list"]
        k, n = so.correct_count(lt, y)
        assert np.float32(k) / np.float32(n) * np.float32(100) == np.float32(case["acc"])
        assert abs(float(so.blend_loss([lt, l2], y)) - case["blend_loss"]) <= 1e-6 * abs(case["blend_loss"])


def test_random_controller_trace():
    gold = json.load(open(os.path.join(G, "random_trace.json")))
    ctl = so.RandomController()
    random.seed(777)
    ctl.on_train_begin()
    it = iter(gold)
    for epoch in range(1, 4):
        ctl.on_epoch_begin(epoch)
        for step in range(6):
            ctl.on_backward_end()
            e, s, cm, car = next(it)
            assert (e, s, cm, car) == (epoch, step, ctl.curation_mode, ctl.caring_modality)


def test_guided_controller_replays_reference_trace():
    """Feed the recorded per-step bucket sums through the restated state machine and
    require the reference's d_BDR / flags (callbacks.py:235-267)."""
    g = json.load(open(os.path.join(G, "guided_trace.json")))
    cfg = g["cfg"]
    ctl = so.GuidedController(cfg["epsilon"], cfg["window"], cfg["starting_epoch"])
    ctl.on_train_begin()
    events = iter(g["trace"])
    pending = None
    n_bdr = 0
    for epoch in range(1, cfg["n_epochs"]):
        ctl.on_epoch_begin(epoch)
        for step in range(1, cfg["train_batches"] + 1):
            ev = next(events)
            if ev["kind"] == "bdr":
                pending = ev
                ev = next(events)
            assert ev["kind"] == "batch" and ev["batch"] == step

            def measure():
                assert pending is not None, "oracle asked for BDR where the reference did not"
                return pending["buckets"]

            before = ctl.d_bdr
            ctl.on_backward_end(measure)
            if pending is not None:
                n_bdr += 1
                assert abs(ctl.d_bdr - pending["d_BDR"]) <= 1e-9 * max(1.0, abs(pending["d_BDR"]))
                pending = None
            else:
                assert ctl.d_bdr == before
            assert float(ctl.curation_mode) == ev["curation_mode"]
            assert ctl.caring_modality == ev["caring_modality"]
            assert abs(ctl.d_bdr - ev["d_BDR"]) <= 1e-9
    assert n_bdr == sum(1 for t in g["trace"] if t["kind"] == "bdr")


def test_bucket_masks_match_reference_counts():
    """main0 = main1 = 62 tensors, bypass0 = bypass1 = 12 (6 shared)."""
    from greedy_multimodal_learning_b200.model import MMTM_MVCNN_names
    names = MMTM_MVCNN_names()
    masks = [so.bucket_mask(n, ["net_view_0", "net_view_1"], ["visual", "skeleton"]) for n in names]
    assert len(names) == 142
    assert sum(1 for m in masks if m & so.BUCKET_MAIN0) == 62
    assert sum(1 for m in masks if m & so.BUCKET_MAIN1) == 62
    assert sum(1 for m in masks if m & so.BUCKET_BYPASS0) == 12
    assert sum(1 for m in masks if m & so.BUCKET_BYPASS1) == 12
    assert sum(1 for m in masks if m == (so.BUCKET_BYPASS0 | so.BUCKET_BYPASS1)) == 6
