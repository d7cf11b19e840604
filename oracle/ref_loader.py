"""Import the UNMODIFIED reference from /root/reference (authoring container only).

TEST INFRASTRUCTURE -- not product code.  Used by `tests/golden/make_golden.py` to
generate the committed golden vectors and by the (skippable) tests that pin the
oracle restatement against the real reference.  `/root/reference` does not exist on
the GPU box; `__graft_entry__.build()` mirrors the unmodified tree into the git-ignored
`baseline/_ref/`, which travels with the snapshot.  Nothing may depend on either being there:
`reference_available()` is the guard.

The reference needs three packages that are not installed in this image
(`gin`, `argh`, `matplotlib`) plus two compat shims for numpy 2 / torch 2.11
(SURVEY.md section 8c).  They are provided here as `sys.modules` stubs; no
reference source is copied or modified.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _default_root():
    """The reference tree where it lies (authoring container), else the git-ignored mirror `baseline/_ref`
    that `__graft_entry__.build()` makes of it (unmodified files; it travels to the GPU box with the snapshot)."""
    for cand in (os.environ.get("GML_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "src", "balanced_mmtm.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "balanced_mmtm.py"))


class _GinRegistry:
    """Just enough of gin-config: `@gin.configurable` injects bound kwargs."""

    def __init__(self):
        self.bindings = {}  # "Name.param" -> value

    def bind(self, name, param, value):
        self.bindings["%s.%s" % (name, param)] = value

    def clear(self):
        self.bindings.clear()


_REGISTRY = _GinRegistry()


def _make_gin_stub():
    gin = types.ModuleType("gin")
    config = types.ModuleType("gin.config")
    config._CONFIG = {}
    config._OPERATIVE_CONFIG = {}

    def configurable(obj=None, **_kw):
        def wrap(target):
            name = target.__name__
            import functools
            import inspect

            if inspect.isclass(target):
                orig_init = target.__init__

                @functools.wraps(orig_init)
                def __init__(self, *a, **k):
                    for key, val in _REGISTRY.bindings.items():
                        n, p = key.split(".")
                        if n == name and p not in k:
                            k[p] = val
                    orig_init(self, *a, **k)

                target.__init__ = __init__
                return target

            @functools.wraps(target)
            def fn(*a, **k):
                for key, val in _REGISTRY.bindings.items():
                    n, p = key.split(".")
                    if n == name and p not in k:
                        k[p] = val
                return target(*a, **k)

            return fn

        if obj is not None and callable(obj):
            return wrap(obj)
        return wrap

    gin.configurable = configurable
    gin.config = config
    gin.parse_config_files_and_bindings = lambda *a, **k: None
    return gin, config


def install_stubs():
    if "gin" not in sys.modules:
        gin, config = _make_gin_stub()
        sys.modules["gin"] = gin
        sys.modules["gin.config"] = config
    if "argh" not in sys.modules:
        argh = types.ModuleType("argh")
        argh.dispatch_command = lambda fn: None
        sys.modules["argh"] = argh
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        style = types.ModuleType("matplotlib.style")
        mpl.style = style
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.style"] = style
    import numpy as np

    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # numpy 2 removed the alias used at callbacks.py:403


def gin_bind(name, param, value):
    _REGISTRY.bind(name, param, value)


def gin_clear():
    _REGISTRY.clear()


@contextlib.contextmanager
def cuda_to_cpu():
    """The reference pins MMTM running stats to 'cuda:0' (balanced_mmtm.py:30-31).
    On a CPU-only host redirect that `.to("cuda:N")` to a no-op while constructing."""
    import torch

    orig_to = torch.Tensor.to

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda") and not torch.cuda.is_available():
            return self
        return orig_to(self, *a, **k)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig_to


def load_reference():
    """Return a namespace with the reference's hot-path modules."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    install_stubs()
    os.environ.setdefault("DATA_DIR", "/nonexistent")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch

    # torch 2.11 dropped ReduceLROnPlateau(verbose=) used at callbacks.py:341-345
    _orig = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(_orig, "_gml_shim", False):

        class _RLROP(_orig):
            _gml_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        torch.optim.lr_scheduler.ReduceLROnPlateau = _RLROP

    ns = types.SimpleNamespace()
    ns.balanced_mmtm = importlib.import_module("src.balanced_mmtm")
    ns.model = importlib.import_module("src.model")
    ns.callbacks = importlib.import_module("src.callbacks")
    ns.framework = importlib.import_module("src.framework")
    ns.training_loop = importlib.import_module("src.training_loop")
    ns.utils = importlib.import_module("src.utils")
    # train.py sets DATA_DIR to a Windows path and imports src.dataset (needs the
    # dataset dir only when called); blend_loss / acc live there (train.py:23-40).
    ns.train = importlib.import_module("train")
    return ns
