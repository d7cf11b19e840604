"""B200-native (sm_100a) hot path of greedy_multimodal_learning.

Public surface = the reference's own names for this path:

    from greedy_multimodal_learning_b200 import MMTM_mitigate, get_rescale_weights      # src/balanced_mmtm.py
    from greedy_multimodal_learning_b200 import Bias_Mitigation_Strong, Bias_Mitigation_Random  # src/callbacks.py
    from greedy_multimodal_learning_b200 import MMTM_MVCNN                                # src/model.py
    from greedy_multimodal_learning_b200 import Model_, blend_loss, acc                   # src/framework.py, train.py

All arithmetic of the path runs in csrc/libgml_b200.so (C ABI: include/gml_b200.h); the
library is loaded on first use and there is no CPU fallback.
"""
from ._lib import GmlError, LIB_PATH, load as load_library  # noqa: F401
from .balanced_mmtm import MMTM_mitigate, SqueezeMeanRecorder, get_mmtm_outputs, get_rescale_weights  # noqa: F401
from .callbacks import Bias_Mitigation_Random, Bias_Mitigation_Strong, Callback, MultiTensorSqnorm  # noqa: F401
from .framework import CallbackList, DevicePrefetcher, Model_, StepIterator, acc, blend_loss  # noqa: F401
from .model import MMTM_MVCNN  # noqa: F401

__version__ = "0.1.0"
