"""Data-parallel hot path on 2 GPUs (NCCL): the MMTM block under batch sharding equals the
single-process block on the concatenated batch, including the global-batch running gate mean and
the curation modes that depend on it; guided training keeps ranks in lock-step.
Run with `gpurun --gpus 2`; skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mmtm_oracle as mo

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(fn, rank, world, port, q):
    import traceback
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    try:
        from greedy_multimodal_learning_b200 import dist as gdist
        gdist.init_from_env("nccl")
        out = fn(rank, world)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, True, out))
    except Exception:
        q.put((rank, False, traceback.format_exc()))


def _spawn(fn, world=2):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    # drain the queue BEFORE joining: a child blocks in put() until its (possibly large) result is read
    results = {}
    import queue as _queue
    import time as _time
    deadline = _time.time() + 300
    while len(results) < world and _time.time() < deadline:
        try:
            r, ok, payload = q.get(timeout=1.0)
            results[r] = (ok, payload)
        except _queue.Empty:
            if all(not p.is_alive() for p in procs) and q.empty():
                break
    for p in procs:
        p.join(30)
        if p.is_alive():
            p.kill()
    assert len(results) == world, "a rank died or hung: %s" % ([p.exitcode for p in procs],)
    for r, (ok, payload) in sorted(results.items()):
        assert ok, "rank %d: %s" % (r, payload)
    return [results[r][1] for r in range(world)]


N, C, H = 6, 32, 8
SEQ = [0, 1, 0, 2, 0]  # normal, curate visual, normal, curate skeleton, normal


def _mmtm_worker(rank, world):
    import greedy_multimodal_learning_b200 as pkg
    from greedy_multimodal_learning_b200 import dist as gdist
    dev = torch.device("cuda", rank)
    p = mo.synth_params(3, C, C)
    m = pkg.MMTM_mitigate(C, C, 4)
    with torch.no_grad():
        for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                             m.fc_skeleton.weight, m.fc_skeleton.bias), p.tensors()):
            dst.copy_(src)
    m.to(dev)
    lo, hi = gdist.shard_batch(N)  # 3 + 3; uneven split below
    if rank == 0:
        lo, hi = 0, 2
    else:
        lo, hi = 2, N
    outs = []
    for i, mode in enumerate(SEQ):
        x = mo.synth_inputs(100 + i, N, C, H)
        kw = {1: dict(curation_mode=True, caring_modality=0), 2: dict(curation_mode=True, caring_modality=1)}.get(mode, {})
        a = x["A"][lo:hi].to(dev).requires_grad_(True)
        b = x["B"][lo:hi].to(dev).requires_grad_(True)
        a_out, b_out, _, _ = m(a, b, **kw)
        torch.autograd.backward([a_out, b_out], [x["gA"][lo:hi].to(dev), x["gB"][lo:hi].to(dev)])
        outs.append(dict(A_out=a_out.detach().cpu().numpy(), dA=a.grad.cpu().numpy(), dB=b.grad.cpu().numpy(),
                         run_v=m.running_avg_weight_visual.cpu().numpy(), lo=lo, hi=hi))
    return outs


def test_mmtm_block_under_batch_sharding_equals_full_batch():
    res = _spawn(_mmtm_worker)
    p = mo.synth_params(3, C, C)
    st = mo.MMTMState.zeros(C)
    for i, mode in enumerate(SEQ):
        x = mo.synth_inputs(100 + i, N, C, H)
        o = mo.forward_backward(x["A"], x["B"], p, st, x["gA"], x["gB"], mode)
        for rank in range(2):
            r = res[rank][i]
            lo, hi = r["lo"], r["hi"]
            np.testing.assert_allclose(r["A_out"], o["A_out"][lo:hi].numpy(), rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(r["dA"], o["dA"][lo:hi].numpy(), rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(r["dB"], o["dB"][lo:hi].numpy(), rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(r["run_v"], st.run_v.numpy(), rtol=1e-5, atol=1e-7)
        assert np.array_equal(res[0][i]["run_v"], res[1][i]["run_v"])  # identical on every rank


def _train_worker(rank, world):
    import greedy_multimodal_learning_b200 as pkg
    from greedy_multimodal_learning_b200 import dist as gdist
    from tests.golden import make_golden_cases as cases
    dev = torch.device("cuda", rank)
    gdist.seed_everything(777)
    model = pkg.MMTM_MVCNN().to(dev)
    model, opt, reducer = gdist.setup_model(model, lambda ps: torch.optim.SGD(ps, lr=0.01))
    cb = pkg.Bias_Mitigation_Strong(0.003, 2, ["net_view_0", "net_view_1"], 1)
    cb.set_model(model, ignore=False)
    flags = []

    class Rec(pkg.Callback):
        def on_batch_end(self, batch, logs):
            flags.append((logs["curation_mode"], logs["caring_modality"], round(logs["d_BDR"], 12)))

    engine = pkg.Model_(model, opt, pkg.blend_loss, 2, metrics=[pkg.acc], data_parallel=reducer).to(dev)
    loader = gdist.ShardedBatches(cases.synth_loader(61, 4, 8, 64))
    engine.train_loop(loader, epochs=2, steps_per_epoch=4, callbacks=[cb, Rec()])
    w = float(model.net_view_0.fc.weight.double().sum())
    return flags, w


def test_guided_training_ranks_stay_in_lock_step():
    (f0, w0), (f1, w1) = _spawn(_train_worker)
    assert f0 == f1            # same statistic, same decisions on every rank, no extra collective
    assert w0 == w1            # replicas bit-identical after all-reduced updates
    assert any(c for c, _, _ in f0)


def _train_cli_worker(rank, world):
    """The reference's guided training command line, one process per GPU (product path, NCCL)."""
    import tempfile
    from greedy_multimodal_learning_b200 import gin_lite
    from greedy_multimodal_learning_b200.train import train
    from tests import frontend_flow as ff
    save = os.path.join(tempfile.gettempdir(), "gml_dp_cli_%s" % os.environ["MASTER_PORT"])
    os.makedirs(save, exist_ok=True)
    gin_lite.clear_config()
    gin_lite.parse_config(ff.TRAIN % dict(
        use_gpu="True", callbacks="['CompletedStopping', 'ReduceLROnPlateau_PyTorch', 'Bias_Mitigation_Strong']"))
    gin_lite.parse_config("train.batch_size=8\nget_mvdcndata.synthetic_samples=(40, 8)")
    H = train(save)
    dist.barrier()
    ckpt = torch.load(os.path.join(save, "model_last_epoch.pt"), map_location="cpu")["model"]
    return ({k: [float(x) for x in H[k]] for k in ("loss", "acc", "val_loss", "test_acc")},
            [sorted(int(i) for i in e) for e in H["train_indices"]], sorted(os.listdir(save)),
            float(sum(v.double().sum() for v in ckpt.values())), torch.cuda.current_device())


def test_train_command_line_one_process_per_gpu():
    (h0, idx0, files0, w0, d0), (h1, idx1, files1, w1, d1) = _spawn(_train_cli_worker)
    assert (d0, d1) == (0, 1)
    assert h0 == h1 and all(np.isfinite(v) for vs in h0.values() for v in vs)
    assert idx0 == idx1 and len(idx0[0]) == 32 and len(set(idx0[0])) == 32
    assert {"history.csv", "history.pickle", "model_best_val.pt", "model_last_epoch.pt"} <= set(files0)
    assert w0 == w1


REC_N, REC_C, REC_H = 10, 32, 8


def _recorder_worker(rank, world):
    """recording.gin under data parallelism: every rank folds the squeezes of ITS shard of the dataset into fp64
    device sums; `result()` all-reduces them (NCCL) and must give the mean over the SELECTED samples of the whole set."""
    import greedy_multimodal_learning_b200 as pkg
    dev = torch.device("cuda", rank)
    p = mo.synth_params(5, REC_C, REC_C)
    blocks = []
    for _ in range(2):
        m = pkg.MMTM_mitigate(REC_C, REC_C, 4)
        with torch.no_grad():
            for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                                 m.fc_skeleton.weight, m.fc_skeleton.bias), p.tensors()):
                dst.copy_(src)
        blocks.append(m.to(dev).eval())
    selected = [0, 2, 3, 7, 9]                      # dataset indices that count (get_rescale_weights' train_indices)
    rec = pkg.SqueezeMeanRecorder(blocks, selected_indices=selected)
    # uneven shards, two batches per rank; rank 1's first batch has no selected sample at all
    shards = {0: [[0, 1, 2], [3, 4]], 1: [[5, 6], [7, 8, 9]]}[rank]
    with torch.no_grad():
        for batch in shards:
            for bi, blk in enumerate(blocks):
                x = mo.synth_inputs(40 + bi, REC_N, REC_C, REC_H)
                blk(x["A"][batch].to(dev), x["B"][batch].to(dev), return_squeezed_mps=True)
            rec.update(batch)
    res = rec.result()
    return [[v.cpu().numpy() for v in blk] for blk in res[1:]]


def test_squeeze_mean_recorder_all_reduces_across_ranks():
    r0, r1 = _spawn(_recorder_worker)
    selected = [0, 2, 3, 7, 9]
    for bi in range(2):
        x = mo.synth_inputs(40 + bi, REC_N, REC_C, REC_H)
        want = [x[k][selected].double().mean(dim=(2, 3)).mean(0).numpy() for k in ("A", "B")]
        for view in range(2):
            assert np.array_equal(r0[bi][view], r1[bi][view])              # rank-identical after the all-reduce
            np.testing.assert_allclose(r0[bi][view], want[view], rtol=1e-6, atol=1e-7)
