"""GPU input pipeline (csrc/prep.cu through hdenseformer_b200.data_utils) against the reference-generated golden vectors
and the numpy/scipy oracle.  Tolerances, written where they apply:
  * crop / flip / one-hot labels / CT window / MR max-normalisation / truncation: bit-exact (same fp32 operations);
  * PET z-score: 2e-6 relative to the channel's range (numpy reduces mean / std pairwise in fp32, the kernel in fp64);
  * warped intensities: 1e-6 of the channel's range (both interpolate in double and round once to fp32; the source
    coordinates come from a BLAS dot in the reference vs explicit FMAs here, ~1e-13 voxels apart);
  * warped labels: bit-exact except voxels whose interpolated indicator is within 1e-6 of the 0.5 threshold (counted)."""
import ast
import os
import random

import numpy as np
import pytest
import torch

from oracle import prep_oracle as PO

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hdenseformer_b200 import data_utils as DU

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "prep_golden.npz"))
CASES = [ast.literal_eval(l) for l in open(os.path.join(HERE, "golden", "prep_golden_cases.txt")) if l.strip()]


def chain_for(kind, patch, ncls, chans, mode, fmode, chain):
    tfs = []
    if "crop" in chain:
        tfs.append(DU.RandomCrop3D(patch))
    if "norm" in chain:
        tfs.append({"petct": DU.PETandCTNormalize, "mr": DU.MRNormalize}[kind]())
    if "warp" in chain:
        tfs.append(DU.RandomTranslationRotationZoom3D(mode=mode, num_class=ncls))
    if "flip" in chain:
        tfs.append(DU.RandomFlip3D(mode=fmode))
    tfs.append(DU.To_Tensor(num_class=ncls, input_channel=chans))
    return DU.Compose(tfs)


def check(gi, gl, ri, rl, kind, warped, chain):
    gi, gl = gi.cpu().numpy(), gl.cpu().numpy()
    assert gi.shape == ri.shape and gl.shape == rl.shape
    for m in range(ri.shape[0]):
        rng = max(float(ri[m].max() - ri[m].min()), 1e-6)
        d = float(np.abs(gi[m] - ri[m]).max())
        exact_channel = not warped and ("norm" not in chain or kind == "mr" or m != 1)
        if exact_channel:
            assert np.array_equal(gi[m], ri[m]), (m, d)
        else:
            assert d <= (1e-6 if warped and not (kind == "petct" and m == 1) else 2e-6) * rng + 1e-7, (m, d, rng)
    mism = int((gl != rl).sum())
    if not warped:
        assert mism == 0
    return mism


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_pipeline_matches_reference_golden(case):
    name, kind, M, vshape, patch, ncls, chans, mode, fmode, chain = case
    img, lab = GOLD[f"{name}__image_in"], GOLD[f"{name}__label_in"]
    seed = int(GOLD[f"{name}__meta"][0])
    random.seed(seed); np.random.seed(seed)
    out = chain_for(kind, patch, ncls, chans, mode, fmode, chain)({"image": img.copy(), "label": lab.copy()})
    assert out["image"].is_cuda and out["image"].dtype == torch.float32
    mism = check(out["image"], out["label"], GOLD[f"{name}__image_out"], GOLD[f"{name}__label_out"], kind, "warp" in chain, chain)
    assert mism == 0, f"{mism} label voxels differ"
    assert torch.equal(out["label"].sum(0), torch.ones_like(out["label"][0]))


@pytest.mark.parametrize("seed", range(6))
def test_pipeline_matches_oracle_random_cases(seed):
    """seeded volumes, all three warp components, 2-4 classes, both normalisations, ragged sizes; resident (CUDA) inputs"""
    rng = np.random.default_rng(seed)
    kind = "petct" if seed % 2 == 0 else "mr"
    M = 2 if kind == "petct" else 3 + seed % 2
    ncls = 2 + seed % 3
    vshape = (int(rng.integers(20, 40)), int(rng.integers(40, 70)), int(rng.integers(40, 70)))
    patch = (min(vshape[0], 24 + 8 * (seed % 2)), 40, 32 + 8 * (seed % 3))
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_prep import synth_volume
    img, lab = synth_volume(rng, M, vshape, ncls, kind)
    mode = ["trz", "tr", "r", "tz", "z", "t"][seed]
    random.seed(seed); np.random.seed(seed)
    origin = PO.draw_crop(img.shape, patch)
    warp_mat = PO.draw_trz(mode)
    flip = PO.draw_flip("hv")
    ri, rl = PO.pipeline(img, lab, patch, ncls, M, norm=kind, origin=origin, warp_mat=warp_mat, flip_axis=flip)
    random.seed(seed); np.random.seed(seed)
    res = DU.ResidentVolumes([{"image": img, "label": lab}])
    out = chain_for(kind, patch, ncls, M, mode, "hv", "crop,norm,warp,flip")(res[0])
    mism = check(out["image"], out["label"], ri, rl, kind, True, "crop,norm,warp,flip")
    # label voxels may only differ where the interpolated indicator sits on the 0.5 threshold
    assert mism <= 2, mism


def test_full_size_properties_and_batching():
    """BASELINE-sized input (2 x 176^3 volume -> 2 x 144^3 patch, batch of 2): identity warp == plain crop bit for bit,
    flip twice == identity, one-hot sums to one, z-scored PET channel, and collate_batch writes the same bits straight
    into preallocated batch tensors."""
    g = torch.Generator(device="cuda").manual_seed(0)
    vol = torch.randn(2, 176, 176, 176, device="cuda", generator=g) * 500
    vol[1] = torch.exp(torch.randn(176, 176, 176, device="cuda", generator=g))
    lab = (torch.rand(176, 176, 176, device="cuda", generator=g) > 0.97).float()
    from hdenseformer_b200 import ops
    size, origin = (144, 144, 144), (7, 19, 30)

    def run(affine, flip):
        io = torch.empty(2, *size, device="cuda"); lo = torch.empty(2, *size, device="cuda")
        aff = None if affine is None else torch.tensor(affine, dtype=torch.float64, device="cuda")
        ops.prep_sample(vol, lab, origin, size, "petct", 0.0, 1024.0, aff, flip, 2, io, lo)
        return io, lo
    base_i, base_l = run(None, 0)
    ident = np.eye(4)[:3].copy()
    wi, wl = run(ident, 0)
    assert torch.equal(wi, base_i) and torch.equal(wl, base_l)
    fi, fl = run(None, 1)
    assert torch.equal(fi.flip(2), base_i) and torch.equal(fl.flip(2), base_l)
    fi, fl = run(None, 2)
    assert torch.equal(fi.flip(3), base_i) and torch.equal(fl.flip(3), base_l)
    assert torch.equal(base_l.sum(0), torch.ones_like(base_l[0]))
    crop = vol[:, 7:151, 19:163, 30:174]
    assert torch.equal(base_i[0], crop[0].clamp(-1024, 1024) / 1024)
    assert abs(base_i[1].double().mean().item()) < 1e-4 and abs(base_i[1].double().std(unbiased=False).item() - 1.0) < 2e-3
    # a pure translation by whole voxels is a shifted crop with zeros where the source leaves the window
    shift = ident.copy(); shift[1, 3] = 3.0
    si, sl = run(shift, 0)
    assert torch.equal(si[:, :, :-3], base_i[:, :, 3:]) and si[:, :, -3:].abs().max().item() == 0
    assert torch.equal(sl[1][:, :-3], base_l[1][:, 3:]) and sl[0][:, -3:].min().item() == 1
    # batching
    ds = DU.DataGenerator(DU.ResidentVolumes([{"image": vol, "label": lab}] * 2), num_class=2,
                          transform=chain_for("petct", size, 2, 2, "tr", "hv", "crop,norm,warp,flip"))
    random.seed(5); np.random.seed(5)
    a = DU.collate_batch(ds, [0, 1])
    random.seed(5); np.random.seed(5)
    bi = torch.empty(2, 2, *size, device="cuda"); bl = torch.empty(2, 2, *size, device="cuda")
    b = DU.collate_batch(ds, [0, 1], bi, bl)
    assert b["image"].data_ptr() == bi.data_ptr()
    assert torch.equal(a["image"], b["image"]) and torch.equal(a["label"], b["label"])
    assert not torch.equal(a["image"][0], a["image"][1])              # two different random draws
