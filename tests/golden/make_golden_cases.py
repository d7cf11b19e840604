"""Case tables and synthetic-input builders shared by make_golden.py (which needs the
reference) and the tests (which do not).  Pure numpy/torch; imports nothing from the reference."""
import numpy as np
import torch

SMALL_CASES = [  # (name, N, C, H, W, seed)
    ("n3c8_5x5", 3, 8, 5, 5, 11),
    ("n2c16_7x7", 2, 16, 7, 7, 12),
    ("n4c12_3x4", 4, 12, 3, 4, 13),
    ("n1c4_1x1", 1, 4, 1, 1, 14),
]
CONFIG_CASES = [("c128_28", 2, 128, 28, 21), ("c256_14", 2, 256, 14, 22), ("c512_7", 2, 512, 7, 23)]
SUB = 97  # stride of the subsample kept for big tensors


SEQUENCE = [  # (mode, N, grad?)   -- mixes train / no_grad eval / curation, batch size varies
    (0, 3, True), (0, 2, False), (1, 3, True), (1, 1, False), (2, 4, True), (0, 3, True), (2, 2, False), (0, 5, True),
]


def synth_history(seed=41, n_total=23, batch=5, dims=(8, 12, 16)):
    """history.pickle pair in the reference's nested-list layout."""
    rs = np.random.RandomState(seed)
    perm = rs.permutation(n_total)
    batches, idx = [], []
    for s in range(0, n_total, batch):
        ids = perm[s:s + batch]
        idx.append(ids)
        batches.append([[torch.from_numpy(rs.standard_normal((len(ids), d)).astype(np.float32)) for _ in range(2)]
                        for d in dims])
    ev = {"test_squeezedmaps_array_list": [batches], "test_indices": [np.concatenate(idx)]}
    sel = np.sort(rs.choice(n_total, size=15, replace=False))
    val = np.array([i for i in range(n_total) if i not in set(sel.tolist())])
    tr = {"train_indices": [sel], "val_indices": [val]}
    return ev, tr


TRACE_CFG = dict(seed=777, image=64, batch=4, train_batches=4, val_batches=1, test_batches=1, n_epochs=4,
                 lr=0.1, epsilon=0.003, window=2, starting_epoch=1, data_seed=61)


def synth_loader(seed, n_batches, batch, image, start=0):
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n_batches):
        x = rs.standard_normal((batch, 2, 3, image, image)).astype(np.float32)
        y = rs.randint(0, 40, size=batch).astype(np.int64)
        idx = np.arange(start + i * batch, start + (i + 1) * batch, dtype=np.int64)
        out.append((torch.from_numpy(idx), torch.from_numpy(x), torch.from_numpy(y)))
    return out

