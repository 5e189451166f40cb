"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + scipy) of the reference's 3-D input pipeline.
Only tests/, __graft_entry__.smoke() and benchmark baselines may import this file; the product path
(hdenseformer_b200/data_utils, csrc/prep.cu) never does.

Each function follows one reference transform and cites it:
  RandomCrop3D                      /root/reference/data_utils/transformer_3d.py:7-42
  RandomTranslationRotationZoom3D   /root/reference/data_utils/transformer_3d.py:45-119
  RandomFlip3D                      /root/reference/data_utils/transformer_3d.py:122-169
  Trunc_and_Normalize / MRNormalize / PETandCTNormalize / To_Tensor
                                    /root/reference/data_utils/data_loader.py:16-36, 39-50, 53-68, 126-159
  chain                             /root/reference/trainer.py:128-141 with config.py:116 transform_3d = [1, 2, 4, 5, 6]

Third-party algorithms the reference calls but this image does not ship (named + restated, as the task allows):
  * skimage.transform.warp(image, coords) (scikit-image, unpinned in the reference's requirements) with a coordinate array:
    scipy.ndimage.map_coordinates(image, coords, order=1 for float input, mode='constant', cval=0, prefilter irrelevant
    for order <= 1) followed by _clip_warp_output: clip to the input's [min, max] except that output samples equal to
    cval keep cval when cval lies outside that range.  scipy IS in this image (1.18) and is what skimage itself calls.
  * transforms3d.euler.euler2mat(a, 0, 0, 'sxyz') = rotation about the first axis by a; transforms3d.affines.compose(T, R,
    Z) = [[R diag(Z), T], [0, 1]].
Parity pinning: tests/golden/prep_*.npz are outputs of the reference's OWN transform classes (imported from
/root/reference by tests/golden/make_golden_prep.py with the three missing third-party modules replaced by the
restatements above), so crop / normalise / flip / one-hot, the control flow and the RNG call order are pinned against the
real code; the warp's interpolation itself rests on scipy.
"""
import random as _pyrandom

import numpy as np
from scipy import ndimage as ndi


# ------------------------------------------------------------------ third-party restatements
def euler2mat_x(angle):
    """transforms3d.euler.euler2mat(angle, 0, 0, 'sxyz')"""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def compose(T, R, Z):
    """transforms3d.affines.compose"""
    A = np.eye(4)
    A[:3, :3] = np.dot(R, np.diag(Z))
    A[:3, 3] = T
    return A


def sk_warp(image, coords):
    """skimage.transform.warp(image, coords) for a float image and a coordinate array (defaults: order=1, mode='constant',
    cval=0, clip=True, preserve_range=False -- float input is not rescaled)."""
    image = np.asarray(image)
    out = ndi.map_coordinates(image, coords, order=1, mode="constant", cval=0.0, prefilter=False)
    mn, mx = image.min(), image.max()
    preserve = not (mn <= 0.0 <= mx)
    if preserve:
        cmask = out == 0.0
    out = np.clip(out, mn, mx)
    if preserve:
        out[cmask] = 0.0
    return out


# ------------------------------------------------------------------ transforms (functional; random draws explicit)
def draw_crop(image_shape, shape, rng=_pyrandom):
    """origins drawn like RandomCrop3D.__call__ (one random.randint per dimension that is larger than the patch)"""
    mm = 1 if len(image_shape) > 3 else 0
    org = [0, 0, 0]
    for i in range(3):
        if image_shape[i + mm] > shape[i]:
            org[i] = rng.randint(0, image_shape[i + mm] - shape[i])
    return tuple(org)


def random_crop3d(image, label, shape, origin):
    """transformer_3d.py:12-42 with the drawn origin; dimensions not larger than the patch are left alone"""
    mm = 1 if image.ndim > 3 else 0
    sl = []
    for i in range(3):
        if image.shape[i + mm] > shape[i]:
            sl.append(slice(origin[i], origin[i] + shape[i]))
        else:
            sl.append(slice(None))
    image = image[(slice(None),) + tuple(sl)] if mm else image[tuple(sl)]
    return image, label[tuple(sl)]


def petct_normalize(image, mean=0, w=1024):
    """data_loader.py:53-68 (in place on a float32 [>=2, D, H, W] array, like the reference)"""
    image[0] = (np.clip(image[0], mean - w, mean + w) - mean) / w
    m = np.mean(image[1])
    s = np.std(image[1])
    image[1] = (image[1] - m) / (s + 1e-3)
    return image


def mr_normalize(image):
    """data_loader.py:39-50"""
    for i in range(image.shape[0]):
        if np.max(image[i]) != 0:
            image[i] = image[i] / np.max(image[i])
    image[image < 0] = 0
    return image


def trunc_and_normalize(image, scale):
    """data_loader.py:16-36"""
    image = image - scale[0]
    rng = scale[1] - scale[0]
    image[image < 0] = 0
    image[image > rng] = rng
    return image / rng


def draw_trz(mode="trz", nprandom=np.random):
    """warp matrix drawn like transformer_3d.py:72-99 (translation draws, then rotation, then zoom)"""
    translation = [0, nprandom.uniform(-5, 5), nprandom.uniform(-5, 5)] if "t" in mode else [0, 0, 0]
    rotation = euler2mat_x(nprandom.uniform(-5, 5) / 180.0 * np.pi) if "r" in mode else euler2mat_x(0.0)
    zoom = [1, nprandom.uniform(0.9, 1.1), nprandom.uniform(0.9, 1.1)] if "z" in mode else [1, 1, 1]
    return compose(translation, rotation, zoom)


def warp_coords(img_size, warp_mat):
    """transformer_3d.py:61-70,101-105"""
    c0, c1, c2 = np.mgrid[:img_size[0], :img_size[1], :img_size[2]]
    coords = np.array([c0 - img_size[0] / 2, c1 - img_size[1] / 2, c2 - img_size[2] / 2])
    tf = np.append(coords.reshape(3, -1), np.ones((1, np.prod(img_size))), axis=0)
    w = np.dot(warp_mat, tf)
    w[0] += img_size[0] / 2
    w[1] += img_size[1] / 2
    w[2] += img_size[2] / 2
    return w[0:3].reshape(3, img_size[0], img_size[1], img_size[2])


def random_trz3d(image, label, warp_mat, num_class):
    """transformer_3d.py:107-117"""
    wc = warp_coords(label.shape, warp_mat)
    if image.ndim > 3:
        for i in range(image.shape[0]):
            image[i] = sk_warp(image[i], wc)
    else:
        image = sk_warp(image, wc)
    new_label = np.zeros(label.shape, dtype=np.float32)
    for z in range(1, num_class):
        temp = sk_warp((label == z).astype(np.float32), wc)
        new_label[temp >= 0.5] = z
    return image, new_label


def draw_flip(mode="hv", nprandom=np.random):
    """1 = flip H (axis -2), 2 = flip W (axis -1), 0 = none -- transformer_3d.py:142-163 ('hv' draws one uniform)"""
    if "h" in mode and "v" in mode:
        return 1 if nprandom.uniform(0, 1) > 0.5 else 2
    if "h" in mode:
        return 1
    if "v" in mode:
        return 2
    return 0


def random_flip3d(image, label, axis):
    if axis == 1:
        image = image[:, :, ::-1, ...] if image.ndim > 3 else image[:, ::-1, ...]
        label = label[:, ::-1, ...]
    elif axis == 2:
        image = image[..., ::-1]
        label = label[..., ::-1]
    return image.copy(), label.copy()


def to_tensor(image, label, num_class, input_channel):
    """data_loader.py:138-152 (numpy arrays; the reference wraps them with torch.from_numpy)"""
    new_image = image[:input_channel, ...] if input_channel > 1 else np.expand_dims(image, axis=0)
    new_label = np.empty((num_class,) + label.shape, dtype=np.float32)
    for z in range(1, num_class):
        new_label[z, ...] = (label == z).astype(np.float32)
    new_label[0, ...] = np.amax(new_label[1:, ...], axis=0) == 0
    return new_image, new_label


def pipeline(image, label, patch, num_class, channels, norm="petct", origin=None, warp_mat=None, flip_axis=0, scale=None):
    """The reference's training chain for one sample with explicit random parameters.  image [M, D, H, W] float32 (copied),
    label [D, H, W] float32."""
    image, label = np.array(image, dtype=np.float32, copy=True), np.array(label, dtype=np.float32, copy=True)
    if origin is not None:
        image, label = random_crop3d(image, label, patch, origin)
        image, label = image.copy(), label.copy()
    if norm == "petct":
        image = petct_normalize(image)
    elif norm == "mr":
        image = mr_normalize(image)
    elif norm == "trunc":
        image = trunc_and_normalize(image, scale)
    if warp_mat is not None:
        image, label = random_trz3d(image, label, warp_mat, num_class)
    image, label = random_flip3d(image, label, flip_axis)
    return to_tensor(image, label, num_class, channels)
