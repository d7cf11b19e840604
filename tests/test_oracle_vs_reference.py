"""Live check of the oracle against the UNMODIFIED reference (only where /root/reference exists,
i.e. in the authoring container; skipped on the GPU box).  Complements the committed goldens."""
import numpy as np
import pytest
import torch

from oracle import mmtm_oracle as mo
from oracle import ref_loader
from oracle import stats_oracle as so
from tests.helpers import assert_close

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load_reference()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(2, 8, 3, 5), (5, 20, 4, 4), (3, 32, 7, 7)])
def test_mmtm_oracle_equals_live_reference(ref, shape, mode):
    n, c, h, w = shape
    x = mo.synth_inputs(n * 31 + c, n, c, h, w)
    p = mo.synth_params(c, c, c)
    with ref_loader.cuda_to_cpu():
        m = ref.balanced_mmtm.MMTM_mitigate(c, c, 4)
    with torch.no_grad():
        for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                             m.fc_skeleton.weight, m.fc_skeleton.bias), p.tensors()):
            dst.copy_(src)
    avg = [0.1 * torch.randn(c), 0.1 * torch.randn(c)]
    kw = {1: dict(curation_mode=True, caring_modality=0), 2: dict(curation_mode=True, caring_modality=1),
          3: dict(turnoff_cross_modal_flow=True, average_squeezemaps=avg)}.get(mode, {})
    st = mo.MMTMState.zeros(c)
    for _ in range(2):  # second call exercises step > 0 in the running mean
        a = x["A"].clone().requires_grad_(True)
        b = x["B"].clone().requires_grad_(True)
        for q in m.parameters():
            q.grad = None
        a_out, b_out, _, _ = m(a, b, **kw)
        torch.autograd.backward([a_out, b_out], [x["gA"], x["gB"]])
        o = mo.forward_backward(x["A"], x["B"], p, st, x["gA"], x["gB"], mode, avg)
        assert_close(o["A_out"], a_out, 1e-6, "A_out")
        assert_close(o["B_out"], b_out, 1e-6, "B_out")
        assert_close(o["dA"], a.grad, 1e-6, "dA")
        assert_close(o["dB"], b.grad, 1e-6, "dB")
        assert_close(o["dWsq"], m.fc_squeeze.weight.grad, 1e-6, "dWsq")
        assert_close(st.run_v, m.running_avg_weight_visual, 1e-6, "run_v")
        assert st.step == m.step


def test_model_mirror_initialises_like_the_reference(ref):
    """Same construction order -> same RNG stream -> identical weights under the reference's seed."""
    import greedy_multimodal_learning_b200 as pkg
    torch.manual_seed(777)
    with ref_loader.cuda_to_cpu():
        r = ref.model.MMTM_MVCNN()
    torch.manual_seed(777)
    mine = pkg.MMTM_MVCNN()
    rs, ms = r.state_dict(), mine.state_dict()
    assert list(rs) == list(ms)
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k


def test_bucket_masks_equal_live_compute_bdr_grouping(ref):
    import greedy_multimodal_learning_b200 as pkg
    names = pkg.model.MMTM_MVCNN_names()
    ref_loader.gin_clear()
    cb = ref.callbacks.Bias_Mitigation_Strong(0.01, 5, ["net_view_0", "net_view_1"], 1)
    for n in names:
        mask = so.bucket_mask(n, cb.branchnames, cb.MMTMnames)
        if "mmtm" in n:
            want = sum(bit for bit, tag in ((so.BUCKET_BYPASS0, "visual"), (so.BUCKET_BYPASS1, "skeleton")) if tag in n)
            want = want or (so.BUCKET_BYPASS0 | so.BUCKET_BYPASS1)
        else:
            want = sum(bit for bit, tag in ((so.BUCKET_MAIN0, "net_view_0"), (so.BUCKET_MAIN1, "net_view_1")) if tag in n)
        assert mask == want, n
