"""Generates tests/golden/prep_*.npz from the reference's OWN transform classes
(/root/reference/data_utils/transformer_3d.py, data_loader.py), imported here with the three third-party modules this
image lacks replaced by the restatements in oracle/prep_oracle.py:
    skimage.transform.warp / resize  -> prep_oracle.sk_warp (scipy.ndimage.map_coordinates) / unused
    transforms3d.euler.euler2mat     -> prep_oracle.euler2mat_x  (only called as euler2mat(a, 0, 0, 'sxyz'))
    transforms3d.affines.compose     -> prep_oracle.compose
    h5py                             -> empty stub (hdf5_reader is not called)
So the crop / normalise / flip / one-hot arithmetic, the control flow and the order of the random draws come from the
reference code itself; the interpolation inside the warp comes from scipy.  Run in the build container only:
    python tests/golden/make_golden_prep.py
"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import prep_oracle as PO  # noqa: E402


def _stub_modules():
    sk = types.ModuleType("skimage"); skt = types.ModuleType("skimage.transform")
    skt.warp = PO.sk_warp
    skt.resize = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("resize is not part of the 3-D chain"))
    sk.transform = skt
    t3 = types.ModuleType("transforms3d"); t3e = types.ModuleType("transforms3d.euler"); t3a = types.ModuleType("transforms3d.affines")

    def euler2mat(ai, aj, ak, axes='sxyz'):
        assert aj == 0 and ak == 0 and axes == 'sxyz'
        return PO.euler2mat_x(ai)
    t3e.euler2mat = euler2mat
    t3a.compose = PO.compose
    sys.modules.update({"skimage": sk, "skimage.transform": skt, "transforms3d": t3, "transforms3d.euler": t3e,
                        "transforms3d.affines": t3a, "h5py": types.ModuleType("h5py")})


def synth_volume(rng, M, shape, num_class, kind):
    D, H, W = shape
    img = np.empty((M, D, H, W), dtype=np.float32)
    if kind == "petct":
        img[0] = rng.normal(0, 600, shape).astype(np.float32)              # HU-like, exceeds the +-1024 window sometimes
        img[1] = np.exp(rng.normal(0, 1, shape)).astype(np.float32)         # heavy-tailed uptake
        img[2:] = rng.normal(0, 1, (M - 2,) + tuple(shape)).astype(np.float32)
    else:
        img[:] = rng.uniform(-0.05, 3.0, (M,) + tuple(shape)).astype(np.float32)
        if M > 2:
            img[2] = 0.0                                                    # an all-zero channel: max == 0 branch
    lab = np.zeros(shape, dtype=np.float32)
    zz, yy, xx = np.mgrid[:D, :H, :W]
    for z in range(1, num_class):
        c = [rng.uniform(0.3, 0.7) * s for s in shape]
        r = [rng.uniform(0.15, 0.3) * s for s in shape]
        lab[((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 <= 1.0] = z
    return img, lab


CASES = [
    # name, kind, M, volume shape, patch, num_class, channels, mode, flip mode, chain
    ("petct_train", "petct", 2, (20, 26, 30), (16, 20, 24), 2, 2, "tr", "hv", "crop,norm,warp,flip"),
    ("petct_val", "petct", 2, (20, 26, 30), (16, 20, 24), 2, 2, None, None, "crop,norm"),
    ("mr_train_trz", "mr", 3, (12, 28, 28), (12, 24, 20), 4, 3, "trz", "hv", "crop,norm,warp,flip"),
    ("petct_flip_only", "petct", 2, (10, 12, 14), (10, 12, 14), 3, 2, None, "h", "norm,flip"),
]


def main():
    _stub_modules()
    sys.path.insert(0, "/root/reference")
    from data_utils import transformer_3d as RT          # the reference's own classes
    from data_utils import data_loader as RD
    out = {}
    for ci, (name, kind, M, vshape, patch, ncls, chans, mode, fmode, chain) in enumerate(CASES):
        rng = np.random.default_rng(100 + ci)
        img, lab = synth_volume(rng, M, vshape, ncls, kind)
        seed = 1234 + ci
        random.seed(seed); np.random.seed(seed)
        tfs = []
        if "crop" in chain:
            tfs.append(RT.RandomCrop3D(patch))
        if "norm" in chain:
            tfs.append(RD.PETandCTNormalize() if kind == "petct" else RD.MRNormalize())
        if "warp" in chain:
            tfs.append(RT.RandomTranslationRotationZoom3D(mode=mode, num_class=ncls))
        if "flip" in chain:
            tfs.append(RT.RandomFlip3D(mode=fmode))
        tfs.append(RD.To_Tensor(num_class=ncls, input_channel=chans))
        sample = {"image": img.copy(), "label": lab.copy()}
        for t in tfs:
            sample = t(sample)
        out[f"{name}__image_in"] = img
        out[f"{name}__label_in"] = lab
        out[f"{name}__image_out"] = np.asarray(sample["image"], dtype=np.float32)
        out[f"{name}__label_out"] = np.asarray(sample["label"], dtype=np.float32)
        out[f"{name}__meta"] = np.array([seed, M, ncls, chans] + list(patch), dtype=np.int64)
        print(name, out[f"{name}__image_out"].shape, out[f"{name}__label_out"].shape,
              "fg voxels", int(out[f"{name}__label_out"][1:].sum()))
    np.savez_compressed(os.path.join(HERE, "prep_golden.npz"), **out)
    with open(os.path.join(HERE, "prep_golden_cases.txt"), "w") as f:
        for c in CASES:
            f.write(repr(c) + "\n")
    print("wrote", os.path.join(HERE, "prep_golden.npz"), os.path.getsize(os.path.join(HERE, "prep_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
