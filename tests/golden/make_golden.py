#!/usr/bin/env python3
"""Generate tests/golden/* by executing the UNMODIFIED reference (authoring container only).

    python tests/golden/make_golden.py          # needs /root/reference

The reference ships no golden vectors or tests, so parity is pinned on
outputs of its own code run here on CPU, torch fp32.  Inputs are NOT stored: they are
re-derived from numpy RandomState seeds by oracle.mmtm_oracle.synth_inputs/synth_params
(stable across numpy versions).  Every array written is an output of reference code:

  mmtm_small.npz         MMTM_mitigate fwd+bwd (balanced_mmtm.py:49-154), tiny shapes, 4 modes, full tensors
  mmtm_config.npz        same at the config shapes 128x28^2 / 256x14^2 / 512x7^2, N=2; compact summaries
  mmtm_sequence.npz      running_avg_* / step over a mixed train/eval/curation call sequence
  rescale.npz            get_rescale_weights (balanced_mmtm.py:179-206) on a synthetic history.pickle pair
  acc.json               train.py:acc / blend_loss incl. the batch-size-2 quirk and argmax ties
  guided_trace.json      3-epoch training_loop run with Bias_Mitigation_Strong on MMTM_MVCNN
                         (64x64 synthetic views): per-step compute_BDR buckets, d_BDR, flags, loss
  random_trace.json      Bias_Mitigation_Random decisions under random.seed(777)
"""
import json
import os
import pickle
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import mmtm_oracle as mo  # noqa: E402

from tests.golden.make_golden_cases import (CONFIG_CASES, SEQUENCE, SMALL_CASES, SUB, TRACE_CFG,  # noqa: E402
                                            synth_history, synth_loader)

torch.set_num_threads(1)  # deterministic reductions in the fixtures
ref = ref_loader.load_reference()
RefMMTM = ref.balanced_mmtm.MMTM_mitigate



def make_ref_module(c, params: mo.MMTMParams):
    with ref_loader.cuda_to_cpu():
        m = RefMMTM(c, c, 4)
    with torch.no_grad():
        m.fc_squeeze.weight.copy_(params.w_sq); m.fc_squeeze.bias.copy_(params.b_sq)
        m.fc_visual.weight.copy_(params.w_v); m.fc_visual.bias.copy_(params.b_v)
        m.fc_skeleton.weight.copy_(params.w_s); m.fc_skeleton.bias.copy_(params.b_s)
    return m


def run_ref(m, x, mode, avg=None, prewarm=None):
    """One reference fwd+bwd in `mode`; returns dict of numpy arrays."""
    if prewarm is not None:  # give the running means a non-trivial history first
        with torch.no_grad():
            m(prewarm["A"], prewarm["B"])
    a = x["A"].clone().requires_grad_(True)
    b = x["B"].clone().requires_grad_(True)
    for p in m.parameters():
        p.grad = None
    kw = {}
    if mode == mo.MODE_XMODAL_OFF:
        kw = dict(turnoff_cross_modal_flow=True, average_squeezemaps=avg)
    elif mode == mo.MODE_CURATE_VISUAL:
        kw = dict(curation_mode=True, caring_modality=0)
    elif mode == mo.MODE_CURATE_SKELETON:
        kw = dict(curation_mode=True, caring_modality=1)
    want_sq = mode != mo.MODE_XMODAL_OFF  # reference raises UnboundLocalError there (:123-124)
    a_out, b_out, scales, sq = m(a, b, True, want_sq, **kw)
    torch.autograd.backward([a_out, b_out], [x["gA"], x["gB"]])
    g = lambda p: (torch.zeros_like(p) if p.grad is None else p.grad).numpy().copy()
    out = dict(A_out=a_out.detach().numpy(), B_out=b_out.detach().numpy(), dA=a.grad.numpy(), dB=b.grad.numpy(),
               gA=scales[0].detach().numpy(), gB=scales[1].detach().numpy(),
               dWsq=g(m.fc_squeeze.weight), dbsq=g(m.fc_squeeze.bias), dWv=g(m.fc_visual.weight),
               dbv=g(m.fc_visual.bias), dWs=g(m.fc_skeleton.weight), dbs=g(m.fc_skeleton.bias),
               wv_has_grad=np.array(m.fc_visual.weight.grad is not None),
               ws_has_grad=np.array(m.fc_skeleton.weight.grad is not None),
               run_v=m.running_avg_weight_visual.numpy().copy(), run_s=m.running_avg_weight_skeleton.numpy().copy(),
               step=np.array(m.step))
    if want_sq:
        out["sA"], out["sB"] = sq[0].detach().numpy(), sq[1].detach().numpy()
    return out


def avg_for(seed, c):
    rs = np.random.RandomState(seed + 1000)
    return [torch.from_numpy((0.1 * rs.standard_normal(c)).astype(np.float32)) for _ in range(2)]


def gen_small():
    store = {}
    for name, n, c, h, w, seed in SMALL_CASES:
        x = mo.synth_inputs(seed, n, c, h, w)
        warm = mo.synth_inputs(seed + 500, n + 1, c, h, w)
        params = mo.synth_params(seed, c, c)
        for mode in range(4):
            m = make_ref_module(c, params)
            r = run_ref(m, x, mode, avg_for(seed, c), prewarm=warm)
            for k, v in r.items():
                store["%s/m%d/%s" % (name, mode, k)] = v
    np.savez_compressed(os.path.join(HERE, "mmtm_small.npz"), **store)
    print("mmtm_small.npz", len(store), "arrays")


def summarize(v):
    flat = v.reshape(-1)
    return dict(sub=flat[::SUB].copy(), sum=np.array(flat.astype(np.float64).sum()),
                abssum=np.array(np.abs(flat.astype(np.float64)).sum()),
                sqsum=np.array((flat.astype(np.float64) ** 2).sum()))


def gen_config():
    store = {}
    for name, n, c, h, seed in CONFIG_CASES:
        x = mo.synth_inputs(seed, n, c, h)
        warm = mo.synth_inputs(seed + 500, 3, c, h)
        params = mo.synth_params(seed, c, c)
        for mode in range(4):
            m = make_ref_module(c, params)
            r = run_ref(m, x, mode, avg_for(seed, c), prewarm=warm)
            for k, v in r.items():
                if v.size > 4096:
                    for kk, vv in summarize(v).items():
                        store["%s/m%d/%s.%s" % (name, mode, k, kk)] = vv
                else:
                    store["%s/m%d/%s" % (name, mode, k)] = v
    np.savez_compressed(os.path.join(HERE, "mmtm_config.npz"), **store)
    print("mmtm_config.npz", len(store), "arrays")




def gen_sequence():
    c, h, seed = 16, 6, 31
    params = mo.synth_params(seed, c, c)
    m = make_ref_module(c, params)
    store = {}
    for i, (mode, n, grad) in enumerate(SEQUENCE):
        x = mo.synth_inputs(seed + 10 * i, n, c, h)
        kw = {}
        if mode == 1:
            kw = dict(curation_mode=True, caring_modality=0)
        if mode == 2:
            kw = dict(curation_mode=True, caring_modality=1)
        if grad:
            a_out, b_out, _, _ = m(x["A"], x["B"], **kw)
        else:
            m.eval()
            with torch.no_grad():
                a_out, b_out, _, _ = m(x["A"], x["B"], **kw)
            m.train()
        store["%d/A_out" % i] = a_out.detach().numpy()
        store["%d/B_out" % i] = b_out.detach().numpy()
        store["%d/run_v" % i] = m.running_avg_weight_visual.numpy().copy()
        store["%d/run_s" % i] = m.running_avg_weight_skeleton.numpy().copy()
        store["%d/step" % i] = np.array(m.step)
    np.savez_compressed(os.path.join(HERE, "mmtm_sequence.npz"), **store)
    print("mmtm_sequence.npz", len(store))




def gen_rescale():
    ev, tr = synth_history()
    store = {}
    with tempfile.TemporaryDirectory() as d:
        e, t = os.path.join(d, "e"), os.path.join(d, "t")
        os.mkdir(e); os.mkdir(t)
        pickle.dump(ev, open(os.path.join(e, "history.pickle"), "wb"))
        pickle.dump(tr, open(os.path.join(t, "history.pickle"), "wb"))
        for validation in (False, True):
            w = ref.balanced_mmtm.get_rescale_weights(e, t, validation=validation, device=torch.device("cpu"))
            assert w[0] is None and len(w) == 4
            for pos in (1, 2, 3):
                for v in (0, 1):
                    store["val%d/pos%d/view%d" % (validation, pos, v)] = w[pos][v].numpy()
    np.savez_compressed(os.path.join(HERE, "rescale.npz"), **store)
    print("rescale.npz", len(store))


def gen_acc():
    rs = np.random.RandomState(51)
    cases = []
    for n in (1, 2, 3, 8):
        logits = rs.standard_normal((n, 5)).astype(np.float32)
        if n == 3:
            logits[1, 2] = logits[1, 4] = logits[1].max() + 1  # exact tie -> first index wins
        y = rs.randint(0, 5, size=n)
        lt, yt = torch.from_numpy(logits), torch.from_numpy(y)
        l2 = torch.from_numpy(rs.standard_normal((n, 5)).astype(np.float32))
        cases.append(dict(n=n, logits=logits.tolist(), logits2=l2.numpy().tolist(), y=y.tolist(),
                          acc=float(ref.train.acc(lt, yt)), acc_list=float(ref.train.acc([lt, l2], yt)),
                          blend_loss=float(ref.train.blend_loss([lt, l2], yt))))
    json.dump(cases, open(os.path.join(HERE, "acc.json"), "w"), indent=1)
    print("acc.json", len(cases))


# ---------------------------------------------------------------- guided training trace


def gen_guided_trace():
    cfg = TRACE_CFG
    ref_loader.gin_clear()
    ref_loader.gin_bind("Bias_Mitigation_Strong", "epsilon", cfg["epsilon"])
    ref_loader.gin_bind("Bias_Mitigation_Strong", "curation_windowsize", cfg["window"])
    ref_loader.gin_bind("Bias_Mitigation_Strong", "starting_epoch", cfg["starting_epoch"])
    ref_loader.gin_bind("Bias_Mitigation_Strong", "branchnames", ["net_view_0", "net_view_1"])
    ref_loader.gin_bind("ProgressionCallback", "other_metrics", [])
    torch.manual_seed(cfg["seed"])
    with ref_loader.cuda_to_cpu():
        model = ref.model.MMTM_MVCNN()
    opt = torch.optim.SGD(model.parameters(), lr=cfg["lr"], weight_decay=0.0, momentum=0)
    cb = ref.callbacks.Bias_Mitigation_Strong()
    trace = []
    orig = cb.compute_BDR

    def spy():
        from oracle.stats_oracle import sqnorm_buckets
        b = sqnorm_buckets(((n, p, p.grad) for n, p in model.named_parameters()), cb.branchnames, cb.MMTMnames)
        d = orig()
        trace.append(dict(kind="bdr", buckets=b, d_BDR=float(d), M=[cb.M_bypass_modal_0, cb.M_bypass_modal_1,
                                                                      cb.M_main_modal_0, cb.M_main_modal_1]))
        return d

    cb.compute_BDR = spy
    orig_end = cb.on_batch_end

    def batch_end(batch, logs):
        orig_end(batch, logs)
        trace.append(dict(kind="batch", batch=batch, loss=logs["loss"], acc=logs["acc"], acc0=logs["acc_modal_0"],
                          acc1=logs["acc_modal_1"], curation_mode=logs["curation_mode"],
                          caring_modality=logs["caring_modality"], d_BDR=logs["d_BDR"]))

    cb.on_batch_end = batch_end
    tr = synth_loader(cfg["data_seed"], cfg["train_batches"], cfg["batch"], cfg["image"])
    va = synth_loader(cfg["data_seed"] + 1, cfg["val_batches"], cfg["batch"], cfg["image"], 1000)
    te = synth_loader(cfg["data_seed"] + 2, cfg["test_batches"], cfg["batch"], cfg["image"], 2000)
    epochs = []
    with tempfile.TemporaryDirectory() as d:
        cwd = os.getcwd()
        os.chdir(d)
        try:
            rec = ref.callbacks.LambdaCallback(on_epoch_end=lambda e, logs: epochs.append(
                {k: (float(v) if isinstance(v, (int, float, np.floating)) else None) for k, v in logs.items()
                 if isinstance(v, (int, float, np.floating)) and k not in ("time", "epoch_begin_time")}))
            ref.training_loop.training_loop(
                model=model, optimizer=opt, loss_function=ref.train.blend_loss, metrics=[ref.train.acc],
                train=tr, valid=va, test=te, steps_per_epoch=len(tr), validation_steps=len(va), test_steps=len(te),
                save_path=d, config={}, custom_callbacks=[cb, rec], use_gpu=False, n_epochs=cfg["n_epochs"],
                nummodalities=2)
        finally:
            os.chdir(cwd)
    final = dict(mmtm_step=[model.mmtm2.step, model.mmtm3.step, model.mmtm4.step],
                 run_v2=model.mmtm2.running_avg_weight_visual.tolist(),
                 fc_w_sum=float(model.net_view_0.fc.weight.double().sum()),
                 mmtm4_wsq_sum=float(model.mmtm4.fc_squeeze.weight.double().sum()))
    json.dump(dict(cfg=cfg, torch=torch.__version__, trace=trace, epochs=epochs, final=final),
              open(os.path.join(HERE, "guided_trace.json"), "w"), indent=1)
    print("guided_trace.json", len(trace), "events;", sum(1 for t in trace if t["kind"] == "bdr"), "BDR calls")
    for t in trace:
        if t["kind"] == "batch":
            print("  batch", t["batch"], "loss %.5f" % t["loss"], "cur", t["curation_mode"], t["caring_modality"],
                  "d_BDR", t["d_BDR"])


def gen_random_trace():
    cb = ref.callbacks.Bias_Mitigation_Random()

    class Holder:
        pass

    h = Holder()
    cb.set_model_pytoune(h)
    random.seed(777)
    cb.on_train_begin({})
    out = []
    for epoch in range(1, 4):
        cb.on_epoch_begin(epoch, {})
        for step in range(6):
            cb.on_backward_end(step)
            out.append([epoch, step, bool(h.curation_mode), h.caring_modality])
    json.dump(out, open(os.path.join(HERE, "random_trace.json"), "w"))
    print("random_trace.json", len(out))


if __name__ == "__main__":
    gen_small()
    gen_config()
    gen_sequence()
    gen_rescale()
    gen_acc()
    gen_random_trace()
    gen_guided_trace()
