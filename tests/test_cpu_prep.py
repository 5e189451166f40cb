"""Input pipeline, CPU side: the oracle restatement against golden vectors produced by the reference's own
transform classes (tests/golden/make_golden_prep.py), and the product's host logic (parameter draws, stage ordering)
against the oracle.  No kernel runs here."""
import ast
import os
import random

import numpy as np
import pytest

from oracle import prep_oracle as PO

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "prep_golden.npz"))
CASES = [ast.literal_eval(l) for l in open(os.path.join(HERE, "golden", "prep_golden_cases.txt")) if l.strip()]


def draw_params(case, img_shape):
    """re-draw the random parameters of a golden case in the order the reference chain consumes them"""
    name, kind, M, vshape, patch, ncls, chans, mode, fmode, chain = case
    seed = int(GOLD[f"{name}__meta"][0])
    random.seed(seed); np.random.seed(seed)
    origin = PO.draw_crop(img_shape, patch) if "crop" in chain else None
    warp_mat = PO.draw_trz(mode) if "warp" in chain else None
    flip = PO.draw_flip(fmode) if "flip" in chain else 0
    return origin, warp_mat, flip


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_golden(case):
    name, kind, M, vshape, patch, ncls, chans, mode, fmode, chain = case
    img, lab = GOLD[f"{name}__image_in"], GOLD[f"{name}__label_in"]
    origin, warp_mat, flip = draw_params(case, img.shape)
    oi, ol = PO.pipeline(img, lab, patch, ncls, chans, norm=("petct" if kind == "petct" else "mr") if "norm" in chain else None,
                         origin=origin, warp_mat=warp_mat, flip_axis=flip)
    assert oi.shape == GOLD[f"{name}__image_out"].shape and ol.shape == GOLD[f"{name}__label_out"].shape
    assert np.array_equal(oi, GOLD[f"{name}__image_out"])          # same numpy arithmetic: bit-exact
    assert np.array_equal(ol, GOLD[f"{name}__label_out"])
    assert np.array_equal(ol.sum(0), np.ones_like(ol[0]))           # one-hot


def test_warp_restatement_edge_semantics():
    """map_coordinates(order=1, mode='constant'): a coordinate outside [0, n-1] gives 0, the last sample is reachable,
    and the clip keeps those zeros when 0 is outside the image's range"""
    a = np.arange(1, 6, dtype=np.float32).reshape(1, 1, 5)
    xs = np.array([-0.5, -1e-9, 0.0, 0.25, 3.5, 4.0, 4.0 + 1e-9, 4.5])
    coords = np.stack([np.zeros_like(xs), np.zeros_like(xs), xs]).reshape(3, 1, 1, -1)
    out = PO.sk_warp(a, coords).ravel()
    assert np.allclose(out, [0, 0, 1, 1.25, 4.5, 5, 0, 0])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_product_transforms_draw_the_reference_parameters(case):
    """the GPU transforms only record parameters when called: same seeds -> same crop origin, warp matrix and flip axis as
    the oracle's draws (i.e. the reference's RNG call order)"""
    from hdenseformer_b200 import data_utils as DU
    from hdenseformer_b200.data_utils._plan import plan_of
    name, kind, M, vshape, patch, ncls, chans, mode, fmode, chain = case
    img, lab = GOLD[f"{name}__image_in"], GOLD[f"{name}__label_in"]
    origin, warp_mat, flip = draw_params(case, img.shape)
    seed = int(GOLD[f"{name}__meta"][0])
    random.seed(seed); np.random.seed(seed)
    sample = {"image": img, "label": lab}
    if "crop" in chain:
        sample = DU.RandomCrop3D(patch)(sample)
    if "norm" in chain:
        sample = (DU.PETandCTNormalize() if kind == "petct" else DU.MRNormalize())(sample)
    if "warp" in chain:
        sample = DU.RandomTranslationRotationZoom3D(mode=mode, num_class=ncls)(sample)
    if "flip" in chain:
        sample = DU.RandomFlip3D(mode=fmode)(sample)
    plan = plan_of(sample)
    if origin is not None:
        assert plan.origin == tuple(origin) and plan.size == tuple(patch)
    if warp_mat is not None:
        assert np.array_equal(plan.affine, warp_mat[:3])
    assert plan.flip_axis == flip
    assert plan.norm == (("petct" if kind == "petct" else "mr") if "norm" in chain else None)


def test_unsupported_orders_fail_loudly():
    from hdenseformer_b200 import data_utils as DU
    img, lab = np.zeros((2, 8, 8, 8), np.float32), np.zeros((8, 8, 8), np.float32)
    s = DU.RandomFlip3D("h")({"image": img, "label": lab})
    with pytest.raises(NotImplementedError):
        DU.RandomCrop3D((4, 4, 4))(s)                        # crop after flip
    s = DU.PETandCTNormalize()({"image": img, "label": lab})
    with pytest.raises(NotImplementedError):
        DU.MRNormalize()(s)                                  # two normalisations
    with pytest.raises(ValueError):
        DU.RandomCrop3D((4, 4, 4))({"image": np.zeros((8, 8), np.float32), "label": lab})


def test_data_generator_roi_selection_and_plan_recording():
    """DataGenerator (data_loader.py:162-210) on in-memory samples: ROI selection as in the reference (single ROI -> binary
    label; list of ROIs -> 1..n), transforms only record (no GPU needed until To_Tensor), samples are not mutated."""
    from hdenseformer_b200 import data_utils as DU
    from hdenseformer_b200.data_utils._plan import plan_of
    rng = np.random.default_rng(0)
    img = rng.normal(size=(2, 12, 16, 20)).astype(np.float32)
    lab = rng.integers(0, 4, size=(12, 16, 20)).astype(np.float32)
    items = [{"image": img, "label": lab}]
    ds = DU.DataGenerator(items, roi_number=2, num_class=2,
                          transform=DU.Compose([DU.RandomCrop3D((8, 16, 10)), DU.PETandCTNormalize(), DU.RandomFlip3D("v")]))
    random.seed(3); np.random.seed(3)
    s = ds[0]
    assert np.array_equal(s["label"], (lab == 2).astype(np.float32))
    p = plan_of(s)
    assert p.size == (8, 16, 10) and p.origin[1] == 0 and 0 <= p.origin[0] <= 4 and 0 <= p.origin[2] <= 10
    assert p.norm == "petct" and (p.p0, p.p1) == (0.0, 1024.0) and p.flip_axis == 2 and p.affine is None
    ds2 = DU.DataGenerator(items, roi_number=[3, 1], num_class=3, transform=None)
    l2 = ds2[0]["label"]
    assert np.array_equal(l2 == 1, lab == 3) and np.array_equal(l2 == 2, lab == 1) and set(np.unique(l2)) <= {0.0, 1.0, 2.0}
    assert np.array_equal(items[0]["label"], lab)               # the stored sample is untouched
    with pytest.raises(ImportError):
        DU.hdf5_reader("/nonexistent.hdf5", "ct")                # h5py is not in this image: loud, not silent
