"""Boundary caller of the hot path: mirror of the reference's `MMTM_MVCNN`
(src/model.py:15-108) -- two torchvision ResNet-18 branches (convolutions stay on
PyTorch/cuDNN, out of scope) with the CUDA MMTM block after layer2/3/4.

Same constructor arguments, parameter names (`net_view_{0,1}.*`, `mmtm{2,3,4}.*`) and
forward signature/return tuple, so reference checkpoints (`{'model': state_dict}`) load
and the reference's `framework.Model_` can drive it unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torchvision.models as models

from . import gin_lite
from .balanced_mmtm import MMTM_mitigate, get_rescale_weights

MMTM_DIMS = ((128, 128, 4), (256, 256, 4), (512, 512, 4))  # src/model.py:58-60


@gin_lite.configurable
class MMTM_MVCNN(nn.Module):
    def __init__(self, nclasses=40, num_views=2, pretraining=False, mmtm_off=False,
                 mmtm_rescale_eval_file_path=None, mmtm_rescale_training_file_path=None, device='cuda:0',
                 saving_mmtm_scales=False, saving_mmtm_squeeze_array=False, mmtm_cls=MMTM_mitigate,
                 mmtm_rescale=None):
        super().__init__()
        if pretraining:
            raise NotImplementedError("no network in this environment: ImageNet weights cannot be fetched")
        self.nclasses, self.num_views = nclasses, num_views
        self.mmtm_off = mmtm_off
        if self.mmtm_off:
            # src/model.py:40-48: dataset-mean squeezes from the recorded history pickles, or
            # handed over directly (e.g. from SqueezeMeanRecorder.result())
            self.mmtm_rescale = mmtm_rescale if mmtm_rescale is not None else get_rescale_weights(
                mmtm_rescale_eval_file_path, mmtm_rescale_training_file_path, validation=False,
                starting_mmtmindice=1, mmtmpositions=4, device=torch.device(device))
        self.saving_mmtm_scales = saving_mmtm_scales
        self.saving_mmtm_squeeze_array = saving_mmtm_squeeze_array
        # construction order fixes the RNG stream -> identical init to the reference under a seed
        self.net_view_0 = models.resnet18(weights=None)
        self.net_view_0.fc = nn.Linear(512, nclasses)
        self.net_view_1 = models.resnet18(weights=None)
        self.net_view_1.fc = nn.Linear(512, nclasses)
        self.mmtm2 = mmtm_cls(*MMTM_DIMS[0])
        self.mmtm3 = mmtm_cls(*MMTM_DIMS[1])
        self.mmtm4 = mmtm_cls(*MMTM_DIMS[2])

    fused_logits_are_view_mean = True  # forward returns (x_0 + x_1) / 2: lets the step engine count on the device

    def mmtm_blocks(self):
        return [self.mmtm2, self.mmtm3, self.mmtm4]

    def forward(self, x, curation_mode=False, caring_modality=None):
        v0, v1 = self.net_view_0, self.net_view_1
        f0 = v0.maxpool(v0.relu(v0.bn1(v0.conv1(x[:, 0, :]))))
        f1 = v1.maxpool(v1.relu(v1.bn1(v1.conv1(x[:, 1, :]))))
        f0, f1 = v0.layer1(f0), v1.layer1(f1)
        scales, squeezed_mps = [], []
        for i in (2, 3, 4):
            f0 = getattr(v0, "layer%d" % i)(f0)
            f1 = getattr(v1, "layer%d" % i)(f1)
            f0, f1, scale, squeezed = getattr(self, "mmtm%d" % i)(
                f0, f1, self.saving_mmtm_scales, self.saving_mmtm_squeeze_array,
                turnoff_cross_modal_flow=bool(self.mmtm_off),
                average_squeezemaps=self.mmtm_rescale[i - 1] if self.mmtm_off else None,
                curation_mode=curation_mode, caring_modality=caring_modality)
            scales.append(scale)
            squeezed_mps.append(squeezed)
        x_0 = v0.fc(torch.flatten(v0.avgpool(f0), 1))
        x_1 = v1.fc(torch.flatten(v1.avgpool(f1), 1))
        return (x_0 + x_1) / 2, [x_0, x_1], scales, squeezed_mps


def MMTM_MVCNN_names():
    """Parameter names of MMTM_MVCNN without allocating weights (meta device)."""
    with torch.device("meta"):
        m = MMTM_MVCNN()
    return [n for n, _ in m.named_parameters()]
