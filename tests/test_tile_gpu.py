"""GPU parity tests of the tile-pipeline kernels (csrc/tile_kernels.cu): one persistent launch per block and
direction -- TMA-fed plane reduction / scaling, the batched FCs on tcgen05 (3xTF32), stages ordered by
release/acquire counters.  Forced with GML_F_FORCE_TILE (fails loudly instead of falling back) and compared with
the CPU oracle (reference src/balanced_mmtm.py:93-154) at the north_star tolerance of 1e-5.
"""
import numpy as np
import pytest
import torch

from greedy_multimodal_learning_b200 import _lib
from oracle import mmtm_oracle as mo
from tests.helpers import assert_close, rel_err
from tests.test_mmtm_gpu import make_module, run_cuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (N, C, H, W): one tile / several tiles / partial last tile, planes of 1, 9, 10, 25, 36, 49, 196, 784 floats
# (scalar and 128-bit plane walks), C = 96 (chunk of 96 planes), every GEMM shape class (one or several column
# tiles, K splits, the two-segment K of dH)
SHAPES = [(1, 32, 4, 4), (2, 32, 7, 7), (4, 32, 1, 1), (3, 32, 5, 2), (33, 32, 3, 3), (5, 64, 6, 6), (3, 96, 5, 5),
          (7, 128, 28, 28), (9, 256, 14, 14), (6, 512, 7, 7), (40, 64, 16, 16), (150, 128, 28, 28),
          (301, 256, 14, 14), (260, 512, 7, 7)]


@pytest.fixture
def tile_tunables(request):
    lib = _lib.load()
    m, lag, gemm, ksplit = request.param
    _lib.check(lib.gml_set_tunable(b"tile_m", m))
    _lib.check(lib.gml_set_tunable(b"tile_lag", lag))
    _lib.check(lib.gml_set_tunable(b"tile_gemm_ctas", gemm))
    _lib.check(lib.gml_set_tunable(b"tile_ksplit_tiles", ksplit))
    yield request.param
    _lib.check(lib.gml_set_tunable(b"tile_m", 0))
    _lib.check(lib.gml_set_tunable(b"tile_lag", 0))
    _lib.check(lib.gml_set_tunable(b"tile_gemm_ctas", 0))
    _lib.check(lib.gml_set_tunable(b"tile_ksplit_tiles", 0))


# (samples per tile, lag, GEMM CTAs, k-tiles per split-K item): automatic; tiny tiles (many tiles -> dependency
# counters, partial-plane ring reuse) with a shallow pipeline and split-K (partials folded in split order by the last
# split to finish); a deep pipeline with a single GEMM CTA (every F item serialised behind its dependencies)
@pytest.mark.parametrize("tile_tunables", [(0, 0, 0, 0), (2, 1, 3, 4), (3, 4, 1, 0), (0, 2, 0, 2)], indirect=True,
                         ids=["auto", "m2_lag1_g3_ks4", "m3_lag4_g1", "auto_lag2_ks2"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "n%dc%d_%dx%d" % s)
def test_tile_pipeline_vs_oracle(shape, tile_tunables):
    n, c, h, w = shape
    if tile_tunables[0] and n > 64:
        pytest.skip("tiny tiles are exercised on the small batches")
    rs = np.random.RandomState(n * 13 + c)
    t = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    x = dict(A=t(n, c, h, w), B=t(n, c, h, w), gA=t(n, c, h, w), gB=t(n, c, h, w))
    warm = dict(A=t(2, c, h, w), B=t(2, c, h, w))
    p = mo.synth_params(c + n, c, c)
    lib = _lib.load()
    m = make_module(c, c, p, _lib.F_FORCE_TILE)
    before = lib.gml_launch_count(6) + lib.gml_launch_count(7)
    r = run_cuda(m, x, 0, None, warm)
    assert lib.gml_launch_count(6) + lib.gml_launch_count(7) == before + 3  # warm fwd + fwd + bwd: one launch each
    st = mo.MMTMState.zeros(c)
    with torch.no_grad():
        mo.forward(warm["A"], warm["B"], p, st, 0)
    o = mo.forward_backward(x["A"], x["B"], p, st, x["gA"], x["gB"], 0)
    for k in ["A_out", "B_out", "dA", "dB", "gA", "gB", "sA", "sB", "dWsq", "dbsq", "dWv", "dbv", "dWs", "dbs"]:
        assert_close(r[k], o[k], 1e-5, k)
    assert_close(r["run_v"], st.run_v, 1e-6, "run_v")
    assert_close(r["run_s"], st.run_s, 1e-6, "run_s")
    # gating identity, bit-exact: one fp32 multiply per element
    assert torch.equal(r["A_out"], x["A"] * r["gA"][:, :, None, None])
    # run-to-run bit reproducibility (K splits are folded in split order, no floating-point atomics)
    m2 = make_module(c, c, p, _lib.F_FORCE_TILE)
    r2 = run_cuda(m2, x, 0, None, warm)
    for k in ["A_out", "B_out", "dA", "dB", "dWsq", "dWv", "dbsq", "gA", "run_v"]:
        assert torch.equal(r[k], r2[k]), k
    # and agreement with the streaming path to fp32 rounding
    m3 = make_module(c, c, p, _lib.F_FORCE_STREAMING)
    r3 = run_cuda(m3, x, 0, None, warm)
    for k in ["A_out", "dA", "dWsq"]:
        assert rel_err(r[k], r3[k]) < 5e-6, k


def test_force_tile_fails_loudly_when_unsupported():
    m = make_module(24, 24, mo.synth_params(1, 24, 24), _lib.F_FORCE_TILE)
    a = torch.randn(2, 24, 5, 5, device=DEV)
    with pytest.raises(_lib.GmlError, match="unsupported"):
        m(a, a)
    m = make_module(32, 32, mo.synth_params(1, 32, 32), _lib.F_FORCE_TILE)
    a = torch.randn(2, 32, 5, 5, device=DEV)
    with pytest.raises(_lib.GmlError, match="unsupported"):  # curation modes take the streaming path
        m(a, a, curation_mode=True, caring_modality=0)
