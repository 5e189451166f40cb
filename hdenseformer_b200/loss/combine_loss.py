"""Drop-in for loss/combine_loss.py: CEPlusDice (:8-35) and DeepSuperloss (:68-79)."""
import torch.nn as nn
import torch.nn.functional as F

from ._fused import seg_loss
from .dice_loss import _check_kwargs


class CEPlusDice(nn.Module):
    """Dice + cross entropy on one-hot float targets; `weight` is a [C] tensor or None; other kwargs go to
    BinaryDiceLoss (smooth)."""

    def __init__(self, weight=None, ignore_index=None, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        self.weight = weight
        self.ignore_index = ignore_index

    def forward(self, predict, target):
        assert predict.size() == target.size()
        return seg_loss([predict], target, self.weight, self.ignore_index, _check_kwargs(self.kwargs))


class DeepSuperloss(nn.Module):
    """sum_i 2^-i * criterion(out_i, nearest-resized target).  With a CEPlusDice criterion all levels run in the
    fused kernels reading the full-resolution target with stride 2^i (no resized copies)."""

    def __init__(self, criterion=None):
        super().__init__()
        self.loss = criterion

    def forward(self, input, target):
        c = self.loss
        if isinstance(c, CEPlusDice):
            full = target.shape[2:]
            levels = []
            nd = len(full)          # 3 (volumes) or 2 (slices of the 2-D model)
            for img in input:
                lv = [full[d] // img.shape[2 + d] for d in range(nd)]
                ok = len(set(lv)) == 1 and lv[0] & (lv[0] - 1) == 0 and all(img.shape[2 + d] * lv[0] == full[d] for d in range(nd))
                levels.append(lv[0].bit_length() - 1 if ok else None)
            if all(l is not None for l in levels):
                return seg_loss(list(input), target, c.weight, c.ignore_index, _check_kwargs(c.kwargs),
                                level_weights=[1 / (2 ** i) for i in range(len(input))], levels=levels)
        loss = 0
        for i, img in enumerate(input):
            label = F.interpolate(target, img.size()[2:])
            loss += self.loss(img, label) * (1 / (2 ** i))
        return loss
