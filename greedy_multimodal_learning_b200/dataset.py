"""Input side of the path: `(idx, data[B,V,3,224,224] fp32, class_id)` batches, the tuple the
step engine unpacks (reference src/dataset.py:95-128, src/framework.py:93-95).

`get_mvdcndata` keeps the reference's gin name and arguments.  With a ModelNet40 directory in the
reference's pre-processed layout (`metadata.json` + `<split>/<model>.npy`, src/dataset.py:97-128)
it reads that; with `synthetic_samples=(n_train, n_test)` (or no `DATA_DIR`) it serves seeded
random images of the same shape, which is all this environment has (no network, no dataset).
Batches come out of pinned memory through `framework.DevicePrefetcher` when `device` is given,
and are sharded per rank under data parallelism (`dist.ShardedBatches`).
"""
from __future__ import annotations

import json
import os
import random
from pathlib import Path

import numpy as np
import torch
import torch.utils.data

from . import gin_lite

_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
_STD = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)


class SyntheticMultiview(torch.utils.data.Dataset):
    """Seeded random views with a class-dependent mean so that a few steps of training move the
    accuracy; item `i` is a pure function of `(seed, i)`."""

    def __init__(self, n, num_views=2, nclasses=40, image_size=224, seed=0):
        self.n, self.num_views, self.nclasses, self.image_size, self.seed = n, num_views, nclasses, image_size, seed

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        class_id = int(torch.randint(0, self.nclasses, (1,), generator=g))
        data = torch.randn(self.num_views, 3, self.image_size, self.image_size, generator=g)
        data += (class_id / self.nclasses - 0.5)
        return idx, data, class_id


class MultiviewModelDataset(torch.utils.data.Dataset):
    """Reference layout: `metadata.json` with `classnames` and per-split sample lists, one tensor file
    of all views per model (src/dataset.py:95-128)."""

    def __init__(self, root_dir, split, num_views=12, specific_view=None, train_flip=False):
        self.root_dir, self.split = Path(root_dir), split
        with open(self.root_dir / 'metadata.json') as f:
            meta = json.load(f)
        self.samples, self.classnames = meta[split], meta['classnames']
        self.specific_view = list(specific_view) if specific_view is not None else list(range(num_views))
        self.train_flip = train_flip

    def __len__(self):
        return len(self.samples)

    def _to_tensor(self, img):
        # HWC uint8 (or already CHW float) -> normalised CHW float, as ToTensor + Normalize do
        t = torch.as_tensor(np.asarray(img))
        if t.ndim == 3 and t.shape[-1] == 3:
            t = t.permute(2, 0, 1)
        t = t.float() / 255.0 if t.dtype == torch.uint8 else t.float()
        if self.train_flip and random.random() < 0.5:
            t = t.flip(-1)
        return (t - _MEAN) / _STD

    def __getitem__(self, idx):
        sample = self.samples[idx]
        class_id = self.classnames.index(sample['classname'])
        imgs = torch.load(self.root_dir / self.split / ('%s.npy' % sample['model']), weights_only=False)
        data = torch.stack([self._to_tensor(imgs[v]) for v in self.specific_view])
        return idx, data, class_id


def _seed(seed, use_cuda):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if use_cuda and torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


@gin_lite.configurable
def get_mvdcndata(ending='.png', root_dir=None, make_npy_files=False, valid_size=0.2, batch_size=8,
                  random_seed_for_validation=10, num_views=12, num_workers=0, specific_views=None, seed=777,
                  use_cuda=True, synthetic_samples=None, image_size=224, nclasses=40, pin_memory=True):
    """-> (training_loader, valid_loader, test_loader); split/shuffle rules of src/dataset.py:68-93."""
    _seed(seed, use_cuda)
    root_dir = root_dir if root_dir is not None else os.environ.get('DATA_DIR')
    views = len(specific_views) if specific_views is not None else num_views
    if synthetic_samples is None and (root_dir is None or not os.path.exists(os.path.join(root_dir, 'metadata.json'))):
        synthetic_samples = (64, 16)
    if synthetic_samples is not None:
        n_train, n_test = synthetic_samples
        training = SyntheticMultiview(n_train, views, nclasses, image_size, seed=seed)
        test_dataset = SyntheticMultiview(n_test, views, nclasses, image_size, seed=seed + 1)
    else:
        training = MultiviewModelDataset(root_dir, 'train', num_views, specific_views, train_flip=True)
        test_dataset = MultiviewModelDataset(root_dir, 'test', num_views, specific_views)
    if not 0 <= valid_size <= 1:
        raise ValueError("valid_size should be in the range [0, 1]")
    indices = list(range(len(training)))
    split = int(np.floor(valid_size * len(training)))
    random.Random(random_seed_for_validation).shuffle(indices)
    training_idx, valid_idx = indices[split:], indices[:split]
    pin = bool(pin_memory and torch.cuda.is_available())
    mk = lambda ds, shuffle: torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=shuffle,
                                                        num_workers=num_workers, pin_memory=pin)
    return (mk(torch.utils.data.Subset(training, training_idx), True),
            mk(torch.utils.data.Subset(training, valid_idx), False), mk(test_dataset, False))
