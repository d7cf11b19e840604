"""Shared comparison helpers.  Tolerances are the ones BASELINE.json's north_star states."""
import numpy as np
import torch

RTOL_MMTM = 1e-5     # MMTM outputs and gradients, fp32, relative
RTOL_STATS = 1e-6    # learning-speed / utilization statistics, relative


def to_np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def rel_err(x, ref):
    """max |x - ref| / max |ref|  (norm-relative, robust to near-zero elements)."""
    x, ref = to_np(x).astype(np.float64), to_np(ref).astype(np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    denom = np.abs(ref).max()
    if denom == 0:
        return float(np.abs(x).max())
    return float(np.abs(x - ref).max() / denom)


def assert_close(x, ref, rtol=RTOL_MMTM, what=""):
    """|x - ref| <= rtol * |ref| + rtol * max|ref| * 0.1 element-wise, and norm-relative <= rtol."""
    e = rel_err(x, ref)
    assert e <= rtol, "%s: norm-relative error %.3e > %.1e" % (what, e, rtol)
    x, ref = to_np(x).astype(np.float64), to_np(ref).astype(np.float64)
    bound = rtol * np.abs(ref) + 0.1 * rtol * np.abs(ref).max()
    bad = np.abs(x - ref) > bound
    assert not bad.any(), "%s: %d elements outside rtol=%g (worst %.3e)" % (
        what, int(bad.sum()), rtol, float((np.abs(x - ref) / np.maximum(np.abs(ref), 1e-30)).max()))
