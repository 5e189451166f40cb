"""Drop-in for loss/cross_entropy.py CrossentropyLoss (:8-22): argmax of the one-hot target, voxel mean."""
import torch.nn as nn

from ._fused import seg_loss


class CrossentropyLoss(nn.Module):
    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight

    def forward(self, inp, target):
        return seg_loss([inp], target, self.weight, None, ce_w=1.0, dice_w=0.0)
