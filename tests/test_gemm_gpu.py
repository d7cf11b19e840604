"""The FC GEMM building block (gml_fc_gemm) against float64 matmul, for every kernel variant the
library can pick: CUDA-core FFMA, mma.sync 3xTF32, tcgen05 (UMMA) 3xTF32.  Shapes cover the FC and
weight-gradient problems of the three blocks, both operand layouts, ragged edges and split-K."""
import numpy as np
import pytest
import torch

from greedy_multimodal_learning_b200 import _lib as L

pytestmark = pytest.mark.gpu

VARIANTS = {"ffma": {"gemm_tf32x3": 0, "gemm_umma": 0}, "mma_sync": {"gemm_tf32x3": 1, "gemm_umma": 0},
            "umma": {"gemm_tf32x3": 1, "gemm_umma": 1}}
DEFAULTS = {"gemm_tf32x3": 1, "gemm_umma": 1, "gemm_big_tiles": 0}

# (m, n, k, a_kc, b_kc): forward FCs are K-major x K-major, dH/dZ are K-major x MN-major, the weight
# gradients MN-major x MN-major
SHAPES = [
    (1024, 512, 1024, 1, 1), (1024, 512, 512, 1, 1), (1024, 1024, 512, 1, 0), (512, 1024, 1024, 0, 0),
    (512, 512, 1024, 0, 0), (256, 512, 1024, 0, 0), (256, 256, 512, 1, 1), (128, 128, 256, 0, 1),
    (1100, 512, 1024, 1, 1), (512, 1024, 1100, 0, 0), (300, 260, 132, 1, 0), (129, 132, 1028, 0, 1),
    (7, 128, 256, 1, 1), (640, 384, 96, 0, 0),
]


def _operand(rs, rows, cols, kc, dev):
    """A(i, k) stored [rows, cols] when kc else [cols, rows]; returns (tensor, ld, logical float64 [rows, cols])."""
    logical = rs.standard_normal((rows, cols)).astype(np.float32)
    stored = logical if kc else np.ascontiguousarray(logical.T)
    t = torch.from_numpy(stored).to(dev)
    return t, stored.shape[1], logical.astype(np.float64)


@pytest.fixture
def tunables():
    lib = L.load()

    def set_(**kw):
        for k, v in kw.items():
            L.check(lib.gml_set_tunable(k.encode(), v))
    yield set_
    set_(**DEFAULTS)


@pytest.mark.parametrize("variant", list(VARIANTS))  # small shapes fall back to the CUDA-core kernel in every variant
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "m%dn%dk%d_%d%d" % s)
def test_fc_gemm_matches_float64(shape, variant, tunables):
    m, n, k, a_kc, b_kc = shape
    lib, dev = L.load(), torch.device("cuda:0")
    tunables(**VARIANTS[variant])
    rs = np.random.RandomState(m * 7 + n * 3 + k)
    a, lda, a64 = _operand(rs, m, k, a_kc, dev)
    b, ldb, b64 = _operand(rs, n, k, b_kc, dev)
    bias = torch.from_numpy(rs.standard_normal(n).astype(np.float32)).to(dev)
    c0 = rs.standard_normal((m, n)).astype(np.float32)
    ws_bytes = lib.gml_fc_gemm_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    prod = a64 @ b64.T
    scale = np.sqrt(k)  # magnitude of one output element
    for act, beta in ((0, 0), (1, 0), (2, 0), (0, 1)):
        c = torch.from_numpy(c0).to(dev)
        L.check(lib.gml_fc_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), bias.data_ptr(), m, n, k, lda, ldb, n, a_kc,
                                b_kc, act, beta, ws.data_ptr(), ws_bytes, st), "gml_fc_gemm")
        want = prod + bias.cpu().numpy().astype(np.float64) + (c0.astype(np.float64) if beta else 0.0)
        if act == 1:
            want = np.maximum(want, 0.0)
        elif act == 2:
            want = 1.0 / (1.0 + np.exp(-want))
        got = c.cpu().numpy().astype(np.float64)
        # fp32-SGEMM accuracy: cuBLAS fp32 (allow_tf32 off) measures 2.0e-4 on these N(0,1) operands at k = 1024
        # (scripts/gemm_accuracy.py); plain TF32 would be ~3e-2
        tol = 1e-5 * scale * (0.25 if act == 2 else 1.0)  # the sigmoid's slope is at most 1/4
        err = np.abs(got - want).max()
        assert err <= tol, "%s act=%d beta=%d: max abs error %.3e > %.3e" % (variant, act, beta, err, tol)
    # the call is repeatable bit for bit (split-K folds in a fixed order; tickets self-reset)
    c1 = torch.empty(m, n, device=dev)
    c2 = torch.empty(m, n, device=dev)
    for out in (c1, c2):
        L.check(lib.gml_fc_gemm(a.data_ptr(), b.data_ptr(), out.data_ptr(), None, m, n, k, lda, ldb, n, a_kc, b_kc, 0, 0,
                                ws.data_ptr(), ws_bytes, st))
    assert torch.equal(c1, c2)


def test_fc_gemm_rejects_bad_arguments():
    lib, dev = L.load(), torch.device("cuda:0")
    x = torch.zeros(64, 64, device=dev)
    p = x.data_ptr()
    assert lib.gml_fc_gemm(None, p, p, None, 64, 64, 64, 64, 64, 64, 1, 1, 0, 0, None, 0, None) == -1
    assert lib.gml_fc_gemm(p, p, p, None, 64, 64, 64, 32, 64, 64, 1, 1, 0, 0, None, 0, None) == -1
    assert lib.gml_fc_gemm(p, p, p, None, 64, 64, 64, 64, 64, 64, 1, 1, 3, 0, None, 0, None) == -1
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    assert lib.gml_fc_gemm(p, p, p, None, 64, 64, 64, 64, 64, 64, 1, 1, 0, 0, ws.data_ptr(), 1024, None) == -3
