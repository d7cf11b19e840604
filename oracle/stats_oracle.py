"""CPU oracle for the analysis statistics of the hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle/mmtm_oracle.py for the rules).

Restates, from scratch:
  * conditional learning speed, `Bias_Mitigation_Strong.compute_BDR`     callbacks.py:199-233
  * the guided rebalancing controller state machine                      callbacks.py:190-197,235-267
  * the random controller                                                callbacks.py:269-302
  * dataset-mean squeezes for conditional utilization, `get_mmtm_outputs`
    + `get_rescale_weights`                                              balanced_mmtm.py:157-206
  * `blend_loss` and `acc` (the bit-exact-count metric)                  train.py:23-40

PARITY STATUS: pinned against the live reference by tests/golden/make_golden.py ->
tests/golden/*.npz|json (see tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
import os
import pickle
import random
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

# bucket bit layout shared with include/gml_b200.h (GML_BUCKET_*)
BUCKET_MAIN0, BUCKET_MAIN1, BUCKET_BYPASS0, BUCKET_BYPASS1 = 1, 2, 4, 8


def bucket_mask(name: str, branchnames: Sequence[str], mmtmnames: Sequence[str]) -> int:
    """Which of the four accumulators a parameter name feeds (callbacks.py:207-223).

    'mmtm' in name -> bypass; inside it every modality tag found in the name gets the
    tensor, and a tensor matching NO tag (fc_squeeze) goes to ALL modalities.  Otherwise
    every branch name contained in the parameter name gets it (a name matching no branch
    is dropped).
    """
    mask = 0
    if "mmtm" in name:
        hit = [i for i, tag in enumerate(mmtmnames) if tag in name]
        for i in (hit if hit else range(len(mmtmnames))):
            mask |= (BUCKET_BYPASS0, BUCKET_BYPASS1)[i]
    else:
        for i, tag in enumerate(branchnames):
            if tag in name:
                mask |= (BUCKET_MAIN0, BUCKET_MAIN1)[i]
    return mask


def sqnorm_buckets(named: Iterable[Tuple[str, torch.Tensor, torch.Tensor]], branchnames, mmtmnames):
    """Per-bucket sums of squares.  `named` yields (name, param, grad).

    Follows callbacks.py:203-205: per tensor an fp32 reduction read back as a Python
    float, then added in double.  Returns dict wn/gn x main/bypass -> [modal0, modal1].
    """
    out = {k: [0.0, 0.0] for k in ("wn_main", "wn_bypass", "gn_main", "gn_bypass")}
    for name, p, g in named:
        wn = (p.detach() ** 2).sum().item()
        gn = (g.detach() ** 2).sum().item()
        m = bucket_mask(name, branchnames, mmtmnames)
        for bit, key, idx in ((BUCKET_MAIN0, "main", 0), (BUCKET_MAIN1, "main", 1),
                              (BUCKET_BYPASS0, "bypass", 0), (BUCKET_BYPASS1, "bypass", 1)):
            if m & bit:
                out["wn_" + key][idx] += wn
                out["gn_" + key][idx] += gn
    return out


class LearningSpeed:
    """Accumulators M_* and d_BDR (callbacks.py:190-197,225-233).  Never reset after
    on_train_begin; divisions and log10 in double."""

    def __init__(self):
        self.m_bypass = [0.0, 0.0]
        self.m_main = [0.0, 0.0]

    def update(self, b: Dict[str, List[float]]) -> float:
        for i in (0, 1):
            self.m_bypass[i] += b["gn_bypass"][i] / b["wn_bypass"][i]
            self.m_main[i] += b["gn_main"][i] / b["wn_main"][i]
        bdr0 = np.log10(self.m_bypass[0] / self.m_main[0])
        bdr1 = np.log10(self.m_bypass[1] / self.m_main[1])
        return float(bdr0 - bdr1)


class GuidedController:
    """Host state machine of Bias_Mitigation_Strong (callbacks.py:235-267).

    `measure` is called exactly when the reference calls compute_BDR and must return the
    sqnorm bucket dict for the CURRENT params/grads.
    """

    def __init__(self, epsilon, curation_windowsize, starting_epoch=2):
        self.epsilon, self.window, self.starting_epoch = epsilon, curation_windowsize, starting_epoch

    def on_train_begin(self):
        self.speed = LearningSpeed()
        self.curation_mode, self.caring_modality = False, None
        self.unlock = False
        self.d_bdr = None

    def on_epoch_begin(self, epoch):
        if epoch >= self.starting_epoch:
            self.unlock = True

    def on_backward_end(self, measure):
        if self.unlock:
            if not self.curation_mode:
                self.d_bdr = self.speed.update(measure())
                if abs(self.d_bdr) > self.epsilon:
                    self.curation_mode = True
                    self.curation_step = 0
                    sign = np.sign(self.d_bdr)
                    if sign == -1:
                        self.caring_modality = 1
                    elif sign == 1:
                        self.caring_modality = 0
                else:
                    self.curation_mode, self.caring_modality = False, 0
            else:
                self.curation_step += 1
                if self.curation_step == self.window:
                    self.curation_mode = False
        else:
            self.d_bdr = self.speed.update(measure())
            self.curation_mode, self.caring_modality = False, 0


class RandomController:
    """Bias_Mitigation_Random (callbacks.py:269-302): global `random`, hard-coded
    starting_epoch 2."""

    starting_epoch = 2

    def on_train_begin(self):
        self.curation_mode, self.caring_modality, self.unlock = False, None, False

    def on_epoch_begin(self, epoch):
        if epoch >= self.starting_epoch:
            self.unlock = True

    def on_backward_end(self, rng=random):
        if self.unlock:
            mode = rng.choice([0, 1, 2])
            self.curation_mode, self.caring_modality = ((False, 0), (True, 1), (True, 0))[mode]
        else:
            self.curation_mode, self.caring_modality = False, 0


# --------------------------------------------------------------------------------------
# conditional utilization inputs
# --------------------------------------------------------------------------------------
def mean_squeezes_from_history(eval_history: dict, train_history: dict, key="test_squeezedmaps_array_list",
                               validation=False, starting_mmtmindice=1, mmtmpositions=4):
    """balanced_mmtm.py:157-206 without the file I/O.

    eval_history[key][0] is a list over batches of [[sA, sB] for each recorded block];
    rows are re-ordered by argsort(eval_history['test_indices'][0]) and averaged over
    train_history['train_indices'][0] (or 'val_indices').  Returns a list of length
    `mmtmpositions`: None below `starting_mmtmindice`, else [mean_view0, mean_view1]
    (float32 numpy).
    """
    n_blocks = mmtmpositions - starting_mmtmindice
    order = np.argsort(eval_history["test_indices"][0])
    per_block = [dict() for _ in range(n_blocks)]
    for batch in eval_history[key][0]:
        assert len(batch) == n_blocks
        for blk, views in enumerate(batch):
            for v, arr in enumerate(views):
                per_block[blk].setdefault(v, []).append(arr.detach().cpu().numpy() if torch.is_tensor(arr) else np.asarray(arr))
    sel = train_history["val_indices" if validation else "train_indices"][0]
    out = []
    for pos in range(mmtmpositions):
        if pos < starting_mmtmindice:
            out.append(None)
            continue
        blk = per_block[pos - starting_mmtmindice]
        out.append([np.concatenate(blk[v])[order][sel].mean(0) for v in sorted(blk)])
    return out


def mean_squeezes_from_files(eval_save_path, training_save_path, **kw):
    with open(os.path.join(eval_save_path, "history.pickle"), "rb") as f:
        ev = pickle.load(f)
    with open(os.path.join(training_save_path, "history.pickle"), "rb") as f:
        tr = pickle.load(f)
    return mean_squeezes_from_history(ev, tr, **kw)


def utilization_rate(acc_multimodal_branch: float, acc_branch_flow_cut: float) -> float:
    """u(m_other | m_i) = (A(y_i) - A(y_i')) / A(y_i): relative accuracy drop of branch i
    when cross-modal flow is cut.  NOT computed anywhere in the reference code
    (formula recalled from the paper, unverified) -- provided for
    convenience only."""
    return (acc_multimodal_branch - acc_branch_flow_cut) / acc_multimodal_branch


# --------------------------------------------------------------------------------------
# loss / metric
# --------------------------------------------------------------------------------------
def blend_loss(y_hat: Sequence[torch.Tensor], y: torch.Tensor) -> torch.Tensor:
    """train.py:23-29 -- sum over views of mean cross entropy."""
    total = 0
    for logits in y_hat:
        total = total + torch.nn.functional.cross_entropy(logits, y)
    return total


def correct_count(y_pred: torch.Tensor, y_true: torch.Tensor) -> Tuple[int, int]:
    """Integer form of train.py:32-40: (number correct, denominator).

    argmax over dim 1 takes the FIRST maximal index on ties (torch .max(1)).  Batch-size-2
    quirk (:36-37): with len(y_true) == 2 every prediction is compared with y_true[0].
    """
    pred = y_pred.max(1)[1]
    tgt = y_true[0] if len(y_true) == 2 else y_true
    return int((pred == tgt).sum().item()), int(pred.numel())


def acc(y_pred, y_true) -> torch.Tensor:
    """train.py:32-40 verbatim semantics, float32 mean * 100."""
    if isinstance(y_pred, list):
        y_pred = torch.mean(torch.stack([o.detach() for o in y_pred], 0), 0)
    pred = y_pred.max(1)[1]
    tgt = y_true[0] if len(y_true) == 2 else y_true
    return (pred == tgt).float().mean() * 100
