"""Drop-in for loss/dice_loss.py (DiceLoss :53-87, BinaryDiceLoss :5-50) on the fused kernels."""
import torch.nn as nn

from ._fused import seg_loss


def _check_kwargs(kwargs):
    smooth = kwargs.get("smooth", 1e-5)
    if kwargs.get("p", 1) != 1 or kwargs.get("reduction", "mean") != "mean":
        raise NotImplementedError("hdenseformer_b200 implements BinaryDiceLoss(p=1, reduction='mean') -- the configuration "
                                  "the reference trainer uses (loss/dice_loss.py:19, trainer.py:763-765)")
    return smooth


class DiceLoss(nn.Module):
    """softmax over C, per-class per-sample soft Dice, batch mean, class mean (loss/dice_loss.py:70-87)."""

    def __init__(self, weight=None, ignore_index=None, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        self.class_weight = weight
        self.ignore_index = ignore_index

    def forward(self, predict, target):
        assert predict.shape == target.shape, 'predict & target shape do not match'
        return seg_loss([predict], target, self.class_weight, self.ignore_index, _check_kwargs(self.kwargs), ce_w=0.0,
                        dice_w=1.0)


class BinaryDiceLoss(nn.Module):
    """Kept for API compatibility (loss/dice_loss.py:5-50): Dice of one probability map, i.e. a 1-channel
    DiceLoss without softmax is not on the hot path; use DiceLoss."""

    def __init__(self, smooth=1e-5, p=1, reduction='mean', k=50):
        super().__init__()
        _check_kwargs(dict(smooth=smooth, p=p, reduction=reduction))
        self.smooth = smooth

    def forward(self, predict, target):
        raise NotImplementedError("BinaryDiceLoss on raw probability maps is only used inside DiceLoss in the reference; "
                                  "call DiceLoss / CEPlusDice")
