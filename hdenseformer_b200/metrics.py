"""On-device metric tail of the train / validation step.

Reference: `compute_dice` / `binary_dice` (trainer.py:891-945) and `RunningDice` (metrics.py:82-151), called every step at
trainer.py:382-398 with `.item()`, `.cpu().numpy()` and an sklearn confusion matrix, i.e. several host synchronisations per
step.  Here ONE kernel (`hdf_confusion_update`, csrc/loss.cu) makes the per-sample confusion counts of
argmax(target) x argmax(logits); everything else is a few tiny device ops on a [B, C, C] tensor.  Nothing synchronises
until the caller reads a value (the reference prints every 10 steps, trainer.py:402-409)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import ops


def batch_confusion(predict: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """[B, C, C] int64 device tensor: conf[b, t, p] = #{voxels of sample b with argmax(target) = t and argmax(predict) = p}.
    `predict` are logits (or probabilities) [B, C, *] fp32 / bf16, `target` is one-hot float of the same shape."""
    if not predict.is_cuda:
        raise RuntimeError("hdenseformer_b200.metrics has no CPU path: tensors must be on a B200 (cuda) device")
    assert predict.shape == target.shape, "predict & target shape do not match"
    ops.ensure_init(predict)
    p = predict.detach()
    if p.dtype not in (torch.float32, torch.bfloat16):
        p = p.float()
    conf = torch.zeros((p.shape[0], p.shape[1], p.shape[1]), dtype=torch.int64, device=p.device)
    return ops.confusion_update(p.contiguous(), target.detach().float().contiguous(), conf)


def dice_from_confusion(conf: torch.Tensor, ignore_index: int = 0, smooth: float = 1e-5) -> torch.Tensor:
    """compute_dice (trainer.py:919-945) from per-sample confusion counts: for every class except `ignore_index` the mean
    over the batch of (2 inter + smooth) / (|pred| + |target| + smooth), rounded to 4 decimals; classes absent from both
    masks count as 1; mean over classes 1.. .  Returns a 0-dim fp32 device tensor (no synchronisation)."""
    m = conf.to(torch.float32)                       # the reference sums 0/1 floats in fp32
    inter = torch.diagonal(m, dim1=1, dim2=2)        # [B, C]
    npred, ntgt = m.sum(1), m.sum(2)
    dice = ((2 * inter + smooth) / (npred + ntgt + smooth)).mean(0)
    dice = torch.round(dice * 1e4) / 1e4
    C = m.shape[1]
    present = (npred.sum(0) + ntgt.sum(0)) > 0
    cls = torch.arange(C, device=m.device)
    dl = torch.where(present & (cls != ignore_index), dice, torch.ones_like(dice))
    return dl[1:].mean() if C > 1 else dl.mean()


def compute_dice_device(predict, target, ignore_index: int = 0, smooth: float = 1e-5) -> torch.Tensor:
    return dice_from_confusion(batch_confusion(predict, target), ignore_index, smooth)


def compute_dice(predict, target, ignore_index: int = 0, smooth: float = 1e-5) -> float:
    """Same value as the reference's compute_dice; one host read instead of one per class."""
    return float(compute_dice_device(predict, target, ignore_index, smooth))


class RunningDice:
    """metrics.py:82-151 with the confusion matrix kept on the device.  `update(logits, onehot_target)` is the fused form of
    the reference's argmax -> cpu -> sklearn.confusion_matrix; `update_matrix(ground_truth, prediction)` keeps the reference's
    signature (integer masks, numpy or tensors)."""

    def __init__(self, labels: Sequence[int], ignore_label: int = 0):
        self.labels = list(labels)
        self.ignore_label = ignore_label
        self.overall_confusion_matrix = None      # [C, C] int64 device tensor (rows = ground truth)

    def _add(self, cm: torch.Tensor):
        # the reference skips a batch whose ground truth is entirely the ignore label (metrics.py:121-123)
        non_ignored = cm.sum() - cm[self.ignore_label].sum() if 0 <= self.ignore_label < cm.shape[0] else cm.sum()
        cm = cm * (non_ignored > 0).to(cm.dtype)
        self.overall_confusion_matrix = cm if self.overall_confusion_matrix is None else self.overall_confusion_matrix + cm

    def update(self, predict: torch.Tensor, target: torch.Tensor):
        conf = batch_confusion(predict, target).sum(0)
        lab = torch.as_tensor(self.labels, device=conf.device)
        self._add(conf.index_select(0, lab).index_select(1, lab))

    def update_matrix(self, ground_truth, prediction):
        gt = torch.as_tensor(ground_truth).reshape(-1).long()
        pr = torch.as_tensor(prediction).reshape(-1).long().to(gt.device)
        n = len(self.labels)
        lut = torch.full((int(max(self.labels)) + 2,), n, dtype=torch.long, device=gt.device)     # unknown labels are dropped
        lut[torch.as_tensor(self.labels, device=gt.device)] = torch.arange(n, device=gt.device)
        gi = lut[gt.clamp(0, lut.numel() - 1)]
        pi = lut[pr.clamp(0, lut.numel() - 1)]
        ok = (gi < n) & (pi < n)
        cm = torch.zeros(n * n + 1, dtype=torch.int64, device=gt.device)
        cm.index_add_(0, torch.where(ok, gi * n + pi, torch.full_like(gi, n * n)), torch.ones_like(gi))
        self._add(cm[:n * n].view(n, n))

    def compute_dice(self, smooth: float = 1e-5) -> Tuple[float, List[float]]:
        cm = self.overall_confusion_matrix.detach().cpu().numpy()      # the only host read
        intersection = np.diag(cm)
        union = cm.sum(axis=1) + cm.sum(axis=0)
        iou = (2 * intersection + smooth) / (union.astype(np.float32) + smooth)
        return float(np.mean(iou[1:])), [round(float(c), 4) for c in iou]

    def init_op(self):
        self.overall_confusion_matrix = None
