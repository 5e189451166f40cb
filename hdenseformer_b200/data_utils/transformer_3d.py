"""3-D augmentation transforms with the reference's names, constructor arguments and random draws
(/root/reference/data_utils/transformer_3d.py:7-169), executed on the GPU.

The reference transforms are numpy functions chained by torchvision's Compose inside DataLoader workers; the affine
warp alone (skimage.transform.warp over a 144^3 patch, once per channel and once per class) costs seconds per sample.
Here a transform does no array work when called: it draws its random parameters exactly like the reference (same RNG,
same number and order of draws, so a seeded run picks the same crops / angles / flips) and records them in the sample's
plan; `To_Tensor` (data_loader.py) then runs the whole chain as one fused gather kernel (csrc/prep.cu) and returns CUDA
tensors.  Supported order = the reference's list order (trainer.py:128-141): crop -> normalise -> warp -> flip ->
To_Tensor, any subset; another order raises NotImplementedError.
"""
import random

import numpy as np

from ._plan import plan_of


class RandomCrop3D(object):
    """transformer_3d.py:7-42: random window of `shape`; dimensions not larger than the patch are kept whole."""

    def __init__(self, shape):
        self.shape = shape
        assert len(self.shape) == 3, 'shape error'

    def __call__(self, sample):
        plan = plan_of(sample)
        dims = plan.size                        # current (D, H, W)
        origin, size = [0, 0, 0], list(dims)
        for i in range(3):
            if dims[i] > self.shape[i]:
                origin[i] = random.randint(0, dims[i] - self.shape[i])     # same draw as the reference
                size[i] = self.shape[i]
        plan.crop(origin, size)
        return sample


def _rot_x(angle):
    """transforms3d.euler.euler2mat(angle, 0, 0, 'sxyz'): rotation about the first (depth) axis"""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


class RandomTranslationRotationZoom3D(object):
    """transformer_3d.py:45-119: in-plane translation (+-5 voxels), rotation about the depth axis (+-5 degrees) and in-plane
    zoom (0.9..1.1), trilinear with zeros outside; labels are warped per class and thresholded at 0.5."""

    def __init__(self, mode='trz', num_class=2):
        self.mode = mode
        self.num_class = num_class

    def __call__(self, sample):
        plan = plan_of(sample)
        u = np.random.uniform
        # same draws, same order as the reference: translation (2 uniforms), rotation (1), zoom (2); absent letters draw nothing
        shift = [0.0, u(-5, 5), u(-5, 5)] if 't' in self.mode else [0.0, 0.0, 0.0]
        angle = u(-5, 5) / 180.0 * np.pi if 'r' in self.mode else 0.0
        scale = [1.0, u(0.9, 1.1), u(0.9, 1.1)] if 'z' in self.mode else [1.0, 1.0, 1.0]
        warp_mat = np.eye(4)                     # transforms3d.affines.compose(T, R, Z) = [[R diag(Z), T], [0, 1]]
        warp_mat[:3, :3] = _rot_x(angle) @ np.diag(scale)
        warp_mat[:3, 3] = shift
        plan.warp(warp_mat, self.num_class)
        return sample


class RandomFlip3D(object):
    """transformer_3d.py:122-169: 'hv' flips H or W (one uniform draw decides which), 'h' / 'v' always flip that axis."""

    def __init__(self, mode='hv'):
        self.mode = mode

    def __call__(self, sample):
        h, v = 'h' in self.mode, 'v' in self.mode
        # 'hv': one uniform decides between the H flip (> 0.5) and the W flip; a single letter always flips its axis
        axis = (1 if np.random.uniform(0, 1) > 0.5 else 2) if (h and v) else (1 if h else 2 if v else 0)
        plan_of(sample).flip(axis)
        return sample
