"""CPU tests: the C-ABI library loads and exports what include/gml_b200.h declares (no
compute calls), the host-side mirrors reproduce the reference's recorded behaviour, and the
product path refuses to run without CUDA."""
import json
import os
import re
import random

import numpy as np
import pytest
import torch

import greedy_multimodal_learning_b200 as pkg
from greedy_multimodal_learning_b200 import _lib, callbacks as cbm, framework as fw
from greedy_multimodal_learning_b200.balanced_mmtm import _mode_from_flags
from oracle import mmtm_oracle as mo
from oracle import stats_oracle as so
from oracle.mmtm_module import OracleMMTM
from tests.golden import make_golden_cases as cases
from tests.helpers import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gml_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gml_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libgml_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header drifted: %s" % (
        declared ^ set(_lib.SIGNATURES))
    assert lib.gml_abi_version() == 1
    assert lib.gml_error_string(-3).decode().startswith("workspace")


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_path_refuses_cpu_tensors():
    m = pkg.MMTM_mitigate(8, 8, 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.randn(2, 8, 3, 3), torch.randn(2, 8, 3, 3))
    sq = pkg.MultiTensorSqnorm([("net_view_0.w", torch.nn.Parameter(torch.ones(3)))], ["net_view_0"], ["visual"])
    with pytest.raises(RuntimeError):
        sq.measure()


def test_constructor_contract():
    m = pkg.MMTM_mitigate(128, 128, 4)
    assert [n for n, _ in m.named_parameters()] == ["fc_squeeze.weight", "fc_squeeze.bias", "fc_visual.weight",
                                                     "fc_visual.bias", "fc_skeleton.weight", "fc_skeleton.bias"]
    assert m.fc_squeeze.weight.shape == (128, 256) and m.fc_visual.weight.shape == (128, 128)
    assert set(m.state_dict()) == {n for n, _ in m.named_parameters()}  # running stats are not persistent
    assert m.step == 0 and m.running_avg_weight_visual.shape == (128,)
    with pytest.raises(NotImplementedError):
        pkg.MMTM_mitigate(8, 8, 4, SEonly=True)
    # identical RNG consumption to the reference's constructor order (fc_squeeze, fc_visual, fc_skeleton)
    torch.manual_seed(3)
    a = pkg.MMTM_mitigate(16, 16, 4)
    torch.manual_seed(3)
    sq, vi, sk = torch.nn.Linear(32, 16), torch.nn.Linear(16, 16), torch.nn.Linear(16, 16)
    assert torch.equal(a.fc_squeeze.weight, sq.weight) and torch.equal(a.fc_skeleton.bias, sk.bias)


def test_mode_flags():
    assert _mode_from_flags(False, None, False) == 0
    assert _mode_from_flags(True, 0, False) == 1
    assert _mode_from_flags(True, 1, False) == 2
    assert _mode_from_flags(False, 0, True) == 3
    with pytest.raises(RuntimeError):
        _mode_from_flags(True, None, False)


def test_bucket_masks_equal_oracle():
    from greedy_multimodal_learning_b200.model import MMTM_MVCNN_names
    for n in MMTM_MVCNN_names() + ["something.else", "mmtm_visual_skeleton.w"]:
        assert cbm.bucket_mask(n, ["net_view_0", "net_view_1"], ["visual", "skeleton"]) == so.bucket_mask(
            n, ["net_view_0", "net_view_1"], ["visual", "skeleton"])


def test_get_rescale_weights_reader_against_reference_golden(tmp_path):
    import pickle
    gold = np.load(os.path.join(G, "rescale.npz"))
    ev, tr = cases.synth_history()
    e, t = tmp_path / "e", tmp_path / "t"
    e.mkdir(); t.mkdir()
    pickle.dump(ev, open(e / "history.pickle", "wb"))
    pickle.dump(tr, open(t / "history.pickle", "wb"))
    for validation in (False, True):
        w = pkg.get_rescale_weights(str(e), str(t), validation=validation, device=torch.device("cpu"))
        assert w[0] is None and len(w) == 4
        for pos in (1, 2, 3):
            for v in (0, 1):
                assert_close(w[pos][v], gold["val%d/pos%d/view%d" % (validation, pos, v)], 1e-6, "rescale")


def test_percent_from_count_and_acc_match_reference_golden():
    for case in json.load(open(os.path.join(G, "acc.json"))):
        lt, y = torch.tensor(case["logits"]), torch.tensor(case["y"])
        k, n = so.correct_count(lt, y)
        assert fw.percent_from_count(k, n) == case["acc"]
        assert float(fw.acc(lt, y)) == case["acc"]
        l2 = torch.tensor(case["logits2"])
        assert float(fw.acc([lt, l2], y)) == case["acc_list"]
        assert abs(float(fw.blend_loss([lt, l2], y)) - case["blend_loss"]) < 1e-6 * abs(case["blend_loss"])


def test_random_controller_matches_reference_trace():
    gold = json.load(open(os.path.join(G, "random_trace.json")))

    class MP:
        pass

    cb, mp = pkg.Bias_Mitigation_Random(), MP()
    cb.set_model_pytoune(mp)
    random.seed(777)
    cb.on_train_begin({})
    it = iter(gold)
    for epoch in range(1, 4):
        cb.on_epoch_begin(epoch, {})
        for step in range(6):
            cb.on_backward_end(step)
            assert next(it) == [epoch, step, bool(mp.curation_mode), mp.caring_modality]


class _CPUStrong(pkg.Bias_Mitigation_Strong):
    """Test double: the 8 bucket sums come from the oracle instead of the CUDA launch."""

    def measure_sqnorms(self):
        return so.sqnorm_buckets(((n, p, p.grad) for n, p in self.model.named_parameters()), self.branchnames,
                                 self.MMTMnames)


def test_host_mirrors_replay_reference_training_trace():
    """MMTM_MVCNN mirror + Model_ mirror + Bias_Mitigation_Strong mirror, with the ORACLE MMTM
    and oracle sqnorm plugged in (CPU), must reproduce the trace recorded from the reference's
    own training_loop (tests/golden/guided_trace.json): losses, accuracies, d_BDR, flags."""
    g = json.load(open(os.path.join(G, "guided_trace.json")))
    if g["torch"] != torch.__version__:
        pytest.skip("golden trace was recorded with torch %s" % g["torch"])
    cfg = g["cfg"]
    torch.manual_seed(cfg["seed"])
    model = pkg.MMTM_MVCNN(mmtm_cls=OracleMMTM)
    opt = torch.optim.SGD(model.parameters(), lr=cfg["lr"], weight_decay=0.0, momentum=0)
    cb = _CPUStrong(cfg["epsilon"], cfg["window"], ["net_view_0", "net_view_1"], cfg["starting_epoch"])
    cb.set_model(model, ignore=False)
    got = []

    class Rec(pkg.Callback):
        def on_batch_end(self, batch, logs):
            got.append(dict(logs))

    engine = pkg.Model_(model, opt, pkg.blend_loss, 2, metrics=[pkg.acc])
    tr = cases.synth_loader(cfg["data_seed"], cfg["train_batches"], cfg["batch"], cfg["image"])
    va = cases.synth_loader(cfg["data_seed"] + 1, cfg["val_batches"], cfg["batch"], cfg["image"], 1000)
    te = cases.synth_loader(cfg["data_seed"] + 2, cfg["test_batches"], cfg["batch"], cfg["image"], 2000)
    torch.set_num_threads(1)
    hist = engine.train_loop(tr, valid_generator=va, test_generator=te, epochs=cfg["n_epochs"] - 1,
                             steps_per_epoch=len(tr), validation_steps=len(va), test_steps=len(te),
                             callbacks=[cb, Rec()])
    want = [t for t in g["trace"] if t["kind"] == "batch"]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a["batch"] == b["batch"]
        assert abs(a["loss"] - b["loss"]) <= 2e-5 * abs(b["loss"]), (a["loss"], b["loss"])
        assert a["acc"] == b["acc"] and a["acc_modal_0"] == b["acc0"] and a["acc_modal_1"] == b["acc1"]
        assert a["curation_mode"] == b["curation_mode"] and a["caring_modality"] == b["caring_modality"]
        assert abs(a["d_BDR"] - b["d_BDR"]) <= 1e-4 * max(1e-2, abs(b["d_BDR"]))
    for h, e in zip(hist, g["epochs"]):
        for k in ("loss", "acc", "val_loss", "val_acc", "test_loss", "test_acc", "acc_modal_0", "val_acc_modal_1"):
            assert abs(h[k] - e[k]) <= 2e-5 * max(1.0, abs(e[k])), k
    assert [m.step for m in model.mmtm_blocks()] == g["final"]["mmtm_step"]
    assert_close(model.mmtm2.running_avg_weight_visual, np.array(g["final"]["run_v2"]), 1e-5, "run_v2")
    assert abs(float(model.mmtm4.fc_squeeze.weight.double().sum()) - g["final"]["mmtm4_wsq_sum"]) < 1e-3


def test_state_dict_keys_match_reference_checkpoint_layout():
    with torch.device("meta"):
        m = pkg.MMTM_MVCNN()
    keys = set(m.state_dict())
    for blk in ("mmtm2", "mmtm3", "mmtm4"):
        for fc in ("fc_squeeze", "fc_visual", "fc_skeleton"):
            assert "%s.%s.weight" % (blk, fc) in keys and "%s.%s.bias" % (blk, fc) in keys
    assert not any("running_avg" in k for k in keys)
    assert sum(p.numel() for p in m.parameters()) == 23_773_008


def test_3xtf32_split_keeps_fp32_accuracy():
    """Arithmetic model of the tensor-core FC GEMMs (gemm_kernels.cu): big = x with the low 13 mantissa bits
    cleared (a TF32 number), small = x - big (exact in fp32; the TF32 datapath keeps its top 19 bits), and
    small_a*big_b + big_a*small_b + big_a*big_b accumulated.  The dropped terms are ~2^-21 relative per
    product -- fp32-SGEMM class -- whereas plain TF32 (big*big only) is ~2^-11."""
    rs = np.random.RandomState(0)
    a = rs.standard_normal((64, 1024)).astype(np.float32)
    b = rs.standard_normal((48, 1024)).astype(np.float32)

    def trunc_tf32(x):
        return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    a_big, b_big = trunc_tf32(a), trunc_tf32(b)
    a_small, b_small = a - a_big, b - b_big
    assert np.array_equal(a_big.astype(np.float64) + a_small.astype(np.float64), a.astype(np.float64))  # exact split
    a_small_t, b_small_t = trunc_tf32(a_small), trunc_tf32(b_small)
    f = lambda x: x.astype(np.float64)
    exact = f(a) @ f(b).T
    three = f(a_small_t) @ f(b_big).T + f(a_big) @ f(b_small_t).T + f(a_big) @ f(b_big).T
    one = f(a_big) @ f(b_big).T
    scale = np.abs(exact).max()
    err3, err1 = np.abs(three - exact).max() / scale, np.abs(one - exact).max() / scale
    assert err3 < 1e-6, err3          # ~3e-7 here: at the fp32 accumulation noise of a length-1024 dot product
    assert err1 > 100 * err3, (err1, err3)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch / CUDA types) and a C
    program must link against the shared library and call it (no GPU needed for these two entry points)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    header_dir = os.path.join(ROOT, "include")
    src = tmp_path / "main.c"
    src.write_text('#include <stdio.h>\n#include "gml_b200.h"\n'
                   "int main(void) {\n"
                   "  gml_mmtm_dims d = {4, 8, 8, 16, 16, 4};\n"
                   '  printf("%d %s %d %d\\n", gml_abi_version(), gml_error_string(GML_E_ALIGN), gml_kernel_tag_count(), '
                   "(int)(gml_mmtm_bwd_workspace_bytes(&d) > 0));\n"
                   "  return 0;\n}\n")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", header_dir, str(src)],
                   check=True)
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "main"
    subprocess.run(["gcc", "-std=c99", "-I", header_dir, str(src), "-o", str(exe), "-L", lib_dir, "-lgml_b200",
                    "-Wl,-rpath," + lib_dir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "1" and int(out[-2]) >= 10 and out[-1] == "1"
