"""ctypes binding of csrc/libgml_b200.so (C ABI: include/gml_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the
product path raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"`
(or `make -C greedy_multimodal_learning_b200/csrc`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgml_b200.so")

MODE_NORMAL, MODE_CURATE_VISUAL, MODE_CURATE_SKELETON, MODE_XMODAL_OFF = 0, 1, 2, 3
F_NO_RUNNING_UPDATE, F_FORCE_STREAMING, F_FORCE_FUSED, F_FORCE_TILE = 1, 2, 4, 8
BUCKET_MAIN0, BUCKET_MAIN1, BUCKET_BYPASS0, BUCKET_BYPASS1 = 1, 2, 4, 8


class GmlError(RuntimeError):
    pass


class MMTMDims(Structure):
    _fields_ = [("n", c_int32), ("c_v", c_int32), ("c_s", c_int32), ("hw_v", c_int32), ("hw_s", c_int32),
                ("d", c_int32)]


_P = c_void_p  # every device/host pointer crosses the boundary as a plain address

# symbol -> (restype, argtypes); mirrors include/gml_b200.h one to one
SIGNATURES = {
    "gml_abi_version": (c_int, []),
    "gml_error_string": (c_char_p, [c_int]),
    "gml_last_cuda_error": (c_char_p, []),
    "gml_device_is_blackwell": (c_int, []),
    "gml_launch_count": (c_int64, [c_int]),
    "gml_kernel_tag_count": (c_int, []),
    "gml_kernel_tag_name": (c_char_p, [c_int]),
    "gml_profile_enable": (None, [c_int]),
    "gml_profile_reset": (None, []),
    "gml_profile_read": (c_int, [c_int, POINTER(c_double), POINTER(c_int64)]),
    "gml_set_tunable": (c_int, [c_char_p, c_int64]),
    "gml_mmtm_fwd_workspace_bytes": (c_size_t, [POINTER(MMTMDims)]),
    "gml_mmtm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P,
                             c_size_t, POINTER(MMTMDims), c_int, c_float, c_uint32, _P]),
    "gml_mmtm_gates": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t,
                               POINTER(MMTMDims), c_int, _P]),
    "gml_mmtm_running": (c_int, [_P, _P, _P, c_int32, c_int64, c_int64, _P]),
    "gml_mmtm_apply": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, POINTER(MMTMDims), c_int, c_float, _P]),
    "gml_mmtm_bwd_workspace_bytes": (c_size_t, [POINTER(MMTMDims)]),
    "gml_mmtm_bwd": (c_int, [_P] * 24 + [c_size_t, POINTER(MMTMDims), c_int, c_float, c_uint32, _P]),
    "gml_sqnorm_workspace_bytes": (c_size_t, [POINTER(c_int64), c_int32]),
    "gml_multi_tensor_sqnorm": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int32), POINTER(c_int32),
                                        c_int32, _P, _P, _P, c_size_t, _P]),
    "gml_squeeze_accumulate": (c_int, [_P, _P, c_int32, c_int32, _P, _P, _P]),
    "gml_accuracy_counts": (c_int, [_P, _P, _P, c_int32, c_int32, _P, _P]),
    "gml_fc_gemm_workspace_bytes": (c_size_t, []),
    "gml_fc_gemm": (c_int, [_P, _P, _P, _P] + [c_int32] * 10 + [_P, c_size_t, _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libgml_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "from the repo root; there is no CPU/PyTorch fallback for the MMTM hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.gml_abi_version() != 1:
        raise ImportError("libgml_b200.so ABI version %d != 1" % lib.gml_abi_version())
    _lib = lib
    return lib


def check(code: int, what: str = ""):
    if code == 0:
        return
    lib = load()
    msg = lib.gml_error_string(code).decode()
    if code == -4:
        msg += ": " + lib.gml_last_cuda_error().decode()
    raise GmlError("%s failed: %s" % (what or "libgml_b200 call", msg))


def ptr(t):
    """Device (or host) address of a torch tensor / None -> NULL."""
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    import torch

    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "greedy_multimodal_learning_b200 runs on CUDA (sm_100a) only; got a %s tensor. "
                "There is no CPU fallback by design." % t.device)
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available; greedy_multimodal_learning_b200 has no CPU path")


def current_stream(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
