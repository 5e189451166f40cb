"""Pins oracle/hdf_oracle.py to the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hdf_oracle as O

torch.set_num_threads(max(1, (os.cpu_count() or 2)))


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name + ".json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["model_nf16_32cube", "model_nf8_aniso", "model2d_nf16_64x48"])
def test_model_forward_backward_matches_reference(golden_dir, name):
    meta, g = _load(golden_dir, name)
    size = tuple(meta["image_size"])
    shapes = O.param_shapes(meta["in_channels"], meta["n_cls"], meta["n_filters"], size, meta["transformer_depth"])
    assert {k: list(v) for k, v in shapes.items()} == meta["shapes"]
    assert list(shapes.keys()) == list(meta["shapes"].keys())
    sd = O.synth_state_dict(shapes, seed=meta["param_seed"])
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    M = meta["in_channels"]
    x = O.synth_petct(meta["batch"], size, seed=meta["data_seed"]) if M == 2 else \
        O.synth_mr(meta["batch"], M, size, seed=meta["data_seed"])
    tgt = O.synth_label(meta["batch"], meta["n_cls"], size, seed=meta["data_seed"])
    outs = O.forward(sd, x, meta["transformer_depth"])
    for i, o in enumerate(outs):
        ref = torch.from_numpy(g[f"out{i}"])
        assert o.shape == ref.shape
        err = (o.detach() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-6, (i, err)
    loss = O.deep_super_loss(outs, tgt, ignore_index=0)
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    for k, (nrm, sm) in meta["grad_stats"].items():
        gn = sd[k].grad.double().norm().item()
        assert abs(gn - nrm) <= 2e-4 * max(nrm, 1e-6) + 1e-7, (k, gn, nrm)
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k])
            got = sd[k[5:]].grad
            den = ref.norm().item() * got.norm().item()
            if ref.norm().item() < 1e-6:
                continue
            cos = (ref.double() * got.double()).sum().item() / den
            assert cos > 0.99999, (k, cos)


def test_loss_cases_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_cases.npz"))
    for tag in ("c3", "c2", "c4"):
        p0 = torch.from_numpy(g[f"{tag}_p"])
        t = torch.from_numpy(g[f"{tag}_t"])
        w = torch.from_numpy(g[f"{tag}_w"])
        for ii, ig in (("ig0", 0), ("ignone", None), ("ig1", 1)):
            for wt, wv in (("w", w), ("nw", None)):
                p = p0.clone().requires_grad_(True)
                l = O.ce_plus_dice(p, t, wv, ig)
                l.backward()
                assert abs(l.item() - float(g[f"{tag}_{ii}_{wt}_loss"])) < 1e-6
                assert np.allclose(p.grad.numpy(), g[f"{tag}_{ii}_{wt}_grad"], rtol=1e-5, atol=1e-9)
        assert abs(O.dice_loss(p0, t, None, 0).item() - float(g[f"{tag}_dice_ig0"])) < 1e-6
        assert abs(O.ce_loss(p0, t).item() - float(g[f"{tag}_ce"])) < 1e-6
    t = torch.from_numpy(g["ds_t"])
    outs = [torch.from_numpy(g[f"ds_p{i}"]).requires_grad_(True) for i in range(4)]
    l = O.deep_super_loss(outs, t, ignore_index=0)
    l.backward()
    assert abs(l.item() - float(g["ds_loss"])) < 1e-6
    for i, o in enumerate(outs):
        assert np.allclose(o.grad.numpy(), g[f"ds_g{i}"], rtol=1e-5, atol=1e-9)


def test_cal_steps_known_values():
    # trainer.py:595-618 restated; 224^3 volume, patch 144, step 72 -> [0,40,80]
    assert O.cal_steps((224, 224, 224), (144,) * 3, (72,) * 3) == [[0, 40, 80]] * 3
    assert O.cal_steps((144, 160, 300), (144,) * 3, (72,) * 3) == [[0], [0, 16], [0, 52, 104, 156]]


def test_sliding_window_identity_net():
    # a "network" that returns fixed logits per voxel: averaging softmax must reproduce argmax
    torch.manual_seed(0)
    vol = torch.randn(2, 40, 36, 50)
    net = lambda d: torch.stack([d[:, 0], d[:, 1]], 1)
    mask, prob = O.sliding_window(net, vol, 2, (32, 32, 32), (16, 16, 16))
    assert mask.shape == (40, 36, 50)
    assert torch.equal(mask, (vol[1] > vol[0]).long())
    assert torch.allclose(prob.sum(1), torch.ones_like(prob.sum(1)), atol=1e-6)
