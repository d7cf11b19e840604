"""Data-parallel host logic on CPU: world_size 2, gloo backend (no GPU needed)."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from greedy_multimodal_learning_b200 import dist as gdist
from greedy_multimodal_learning_b200.balanced_mmtm import _MMTMFunction


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(fn, world=2):
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
    deadline = _time.time() + 170
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


def _entry(fn, rank, world, port, q):
    import traceback
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    try:
        gdist.init_from_env("gloo")
        out = fn(rank, world)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, True, out))
    except Exception:
        q.put((rank, False, traceback.format_exc()))


def _tiny_model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                               torch.nn.Linear(16, 3))


def _data():
    rs = np.random.RandomState(0)
    return torch.from_numpy(rs.standard_normal((10, 6)).astype(np.float32)), torch.from_numpy(rs.randint(0, 3, 10))


def _grad_worker(rank, world):
    x, y = _data()
    model = _tiny_model()
    gdist.broadcast_parameters(model)
    red = gdist.GradientAllReduce(model, bucket_mb=0.0005)  # tiny buckets -> several collectives
    assert len(red.buckets) >= 3
    opt = gdist.DPOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), red)
    lo, hi = gdist.shard_batch(len(y))
    out = []
    for step in range(2):
        opt.zero_grad()
        # sum-reduced loss scaled by the GLOBAL batch -> averaging over ranks must not be applied twice
        loss = torch.nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi], reduction="sum") / len(y) * world
        loss.backward()
        red.finish()
        out.append([p.grad.clone() for p in model.parameters()])
        opt.step()
    # a parameter that receives no gradient (substituted excitation FC during a curation window)
    opt.zero_grad()
    h = model[0](x[lo:hi]).relu()
    h.sum().backward()
    red.finish()  # must not hang although most hooks never fired
    # no gradient this step -> .grad is None for the optimizer (single-GPU zero_grad(set_to_none=True) semantics)
    assert model[4].weight.grad is None and model[2].weight.grad is None and model[0].weight.grad is not None
    opt.zero_grad()
    assert model[4].weight.grad is not None and float(model[4].weight.grad.abs().sum()) == 0.0  # view re-attached
    return [[g.numpy() for g in grads] for grads in out], [p.detach().numpy() for p in model.parameters()]


def test_gradient_allreduce_matches_single_process_full_batch():
    res = _spawn(_grad_worker)
    x, y = _data()
    model = _tiny_model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for step in range(2):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y, reduction="sum").div(len(y)).backward()
        want = [p.grad.numpy().copy() for p in model.parameters()]
        for rank in range(2):
            for g, w in zip(res[rank][0][step], want):
                np.testing.assert_allclose(g, w, rtol=1e-5, atol=1e-7)
        opt.step()
    for a, b in zip(res[0][1], res[1][1]):  # replicas stay identical
        assert np.array_equal(a, b)


def _window_worker(rank, world):
    """Momentum + weight decay across steps in which the last layer gets NO gradient (what a curation window does to
    the substituted side's excitation FC): data-parallel weights must follow the single-process run."""
    x, y = _data()
    model = _tiny_model()
    gdist.broadcast_parameters(model)
    red = gdist.GradientAllReduce(model, bucket_mb=0.0005)
    opt = gdist.DPOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=0.01), red)
    lo, hi = gdist.shard_batch(len(y))
    for step in range(6):
        opt.zero_grad()
        if step in (2, 3):   # "window": the head is bypassed
            loss = model[2](model[1](model[0](x[lo:hi]))).pow(2).mean()
        else:
            loss = torch.nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi])
        loss.backward()
        red.finish()
        opt.step()
    return [p.detach().numpy() for p in model.parameters()]


def test_data_parallel_matches_single_process_across_a_no_gradient_window():
    res = _spawn(_window_worker)
    x, y = _data()
    model = _tiny_model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=0.01)
    for step in range(6):
        opt.zero_grad()  # set_to_none=True: the head's .grad is None inside the window -> SGD skips it
        if step in (2, 3):
            loss = model[2](model[1](model[0](x))).pow(2).mean()
        else:
            loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
    for rank in range(2):
        for got, p in zip(res[rank], model.parameters()):
            np.testing.assert_allclose(got, p.detach().numpy(), rtol=2e-5, atol=1e-6)


def test_sharded_batches_remainder_policy():
    mk = lambda n: [(torch.arange(n), torch.arange(n * 2).view(n, 2).float(), torch.arange(n) % 3)]
    # 7 samples on 2 ranks: drop the last one -> 3 + 3, identical sizes, no empty shard
    a = list(gdist.ShardedBatches(mk(7), 0, 2))[0]
    b = list(gdist.ShardedBatches(mk(7), 1, 2))[0]
    assert len(a[2]) == len(b[2]) == 3 and torch.equal(torch.cat([a[0], b[0]]), torch.arange(6))
    # fewer samples than ranks: the batch is skipped on every rank
    assert list(gdist.ShardedBatches(mk(1), 0, 2)) == [] and list(gdist.ShardedBatches(mk(1), 1, 2)) == []
    # pad: 7 -> 8 by repeating the first sample
    a = list(gdist.ShardedBatches(mk(7), 0, 2, remainder="pad"))[0]
    b = list(gdist.ShardedBatches(mk(7), 1, 2, remainder="pad"))[0]
    assert len(a[2]) == len(b[2]) == 4 and int(b[0][-1]) == 0
    with pytest.raises(ValueError):
        gdist.ShardedBatches(mk(4), 0, 2, remainder="keep")


def test_shard_batch_partitions_exactly():
    for n in (1, 2, 7, 8, 2048):
        for world in (1, 2, 3, 8):
            spans = [gdist.shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    loader = [(torch.arange(8), torch.arange(8 * 2).view(8, 2), torch.arange(8) % 3)]
    s0 = list(gdist.ShardedBatches(loader, 0, 2))[0]
    s1 = list(gdist.ShardedBatches(loader, 1, 2))[0]
    assert torch.equal(torch.cat([s0[0], s1[0]]), torch.arange(8)) and s0[1].shape == (4, 2)


def _running_worker(rank, world):
    # per-rank gate sums with DIFFERENT local batch sizes; the all-reduce carries [sum | n]
    rs = np.random.RandomState(3)
    gates = torch.from_numpy(rs.uniform(0, 1, (7, 5)).astype(np.float32))
    local = gates[:3] if rank == 0 else gates[3:]
    buf = torch.cat([local.sum(0), torch.tensor([float(local.shape[0])])])
    dist.all_reduce(buf)
    run_v = torch.full((5,), 0.25)
    run_s = torch.full((5,), 0.25)
    _MMTMFunction._running_update_dp(run_v, run_s, buf[:5], buf[5], 3)
    return run_v.numpy(), gates.numpy()


def test_running_gate_mean_is_the_global_batch_mean():
    (rv0, gates), (rv1, _) = _spawn(_running_worker)
    want = (gates.mean(0) + 0.25 * 3) / 4  # balanced_mmtm.py:113 on the concatenated batch
    np.testing.assert_allclose(rv0, want, rtol=1e-6)
    assert np.array_equal(rv0, rv1)


def _random_ctl_worker(rank, world):
    from greedy_multimodal_learning_b200 import Bias_Mitigation_Random

    class MP:
        pass

    gdist.seed_everything(777)
    cb, mp_ = Bias_Mitigation_Random(), MP()
    cb.set_model_pytoune(mp_)
    cb.on_train_begin({})
    out = []
    for epoch in (1, 2, 3):
        cb.on_epoch_begin(epoch, {})
        for step in range(5):
            cb.on_backward_end(step)
            out.append((bool(mp_.curation_mode), mp_.caring_modality))
    return out


def test_random_controller_is_rank_identical_under_shared_seed():
    a, b = _spawn(_random_ctl_worker)
    assert a == b and any(m for m, _ in a)


def _train_cli_worker(rank, world):
    """`train` entry point under a 2-process launch, oracle block plugged in (CPU has no product path)."""
    import tempfile
    from greedy_multimodal_learning_b200 import gin_lite
    from greedy_multimodal_learning_b200.train import train
    from oracle.mmtm_module import OracleMMTM
    from tests import frontend_flow as ff
    save = os.path.join(tempfile.gettempdir(), "gml_dp_cli_%s" % os.environ["MASTER_PORT"])
    os.makedirs(save, exist_ok=True)
    gin_lite.clear_config()
    gin_lite.parse_config(ff.TRAIN % dict(use_gpu="False", callbacks="['CompletedStopping', 'ReduceLROnPlateau_PyTorch']"))
    gin_lite.parse_config("training_loop.n_epochs=3\nget_mvdcndata.synthetic_samples=(20, 4)\n"
                          "get_mvdcndata.image_size=32")
    gin_lite.bind_parameter("MMTM_MVCNN.mmtm_cls", OracleMMTM)
    H = train(save)
    dist.barrier()
    files = sorted(os.listdir(save))
    ckpt = torch.load(os.path.join(save, "model_last_epoch.pt"), map_location="cpu")["model"]
    return ({k: [float(x) for x in H[k]] for k in ("loss", "acc", "val_loss", "test_acc")},
            [sorted(int(i) for i in e) for e in H["train_indices"]], files,
            float(sum(v.double().sum() for v in ckpt.values())))


def test_train_command_line_under_two_ranks():
    (h0, idx0, files0, w0), (h1, idx1, files1, w1) = _spawn(_train_cli_worker)
    assert h0 == h1                                   # epoch sums are all-reduced -> identical logs on every rank
    assert idx0 == idx1 and len(idx0[0]) == 16 and len(set(idx0[0])) == 16  # 20 samples, valid_size 0.2
    assert {"history.csv", "history.pickle", "model_best_val.pt", "model_last_epoch.pt"} <= set(files0)
    assert w0 == w1 and np.isfinite(w0)
