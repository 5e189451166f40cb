"""HDenseFormer_2D (reference models/HDenseFormer_2D.py) on the B200 path: the 2-D graph run as flat volumes on
the 3-D engine and kernels, against the golden vectors produced by the reference's own 2-D module and against the live oracle
(oracle/hdf_oracle.py, pinned bit-exactly to that module).  Same gates as the 3-D model: fp32 logits 1e-4 with an exact
argmax mask, loss 1e-4, per-tensor gradient cosine >= 0.999; bf16 logits 3e-2 (small 2-D slices carry fewer voxels per
InstanceNorm statistic than the 3-D volumes: 2e-2 is met at 384^2)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hdf_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
    from hdenseformer_b200.models import HDenseFormer_2D

DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
# conv biases in front of a non-affine InstanceNorm: their gradient is analytically zero (rounding noise in both implementations)
ZERO_GRAD_KEYS = ("deep_conv.double_conv.0.bias", "up1.double_conv.0.bias", "up2.double_conv.0.bias", "up3.double_conv.0.bias")


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().cpu().abs().max().clamp_min(1e-12)).item()


def test_fp32_forward_backward_matches_reference_golden_2d():
    meta = json.load(open(os.path.join(HERE, "golden", "model2d_nf16_64x48.json")))
    g = np.load(os.path.join(HERE, "golden", "model2d_nf16_64x48.npz"))
    size = tuple(meta["image_size"])
    shapes = O.param_shapes(meta["in_channels"], meta["n_cls"], meta["n_filters"], size, meta["transformer_depth"])
    sd = O.synth_state_dict(shapes, seed=meta["param_seed"])
    m = HDenseFormer_2D(meta["in_channels"], meta["n_cls"], meta["n_filters"], size, meta["transformer_depth"])
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x = O.synth_mr(meta["batch"], meta["in_channels"], size, seed=meta["data_seed"])
    tgt = O.synth_label(meta["batch"], meta["n_cls"], size, seed=meta["data_seed"])
    outs = m(x.to(DEV))
    assert [tuple(o.shape) for o in outs] == [tuple(g[f"out{i}"].shape) for i in range(4)]
    for i, o in enumerate(outs):
        assert rel(o, torch.from_numpy(g[f"out{i}"])) < 1e-4, i
    assert (outs[0].argmax(1).cpu() != torch.from_numpy(g["out0"]).argmax(1)).sum().item() == 0
    loss = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))(outs, tgt.to(DEV))
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    for k, (nrm, _) in meta["grad_stats"].items():
        if nrm < 1e-6:
            continue                      # conv biases in front of a non-affine InstanceNorm: analytically zero
        gn = grads[k].double().norm().item()
        assert abs(gn - nrm) <= 2e-3 * nrm + 1e-7, (k, gn, nrm)
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k]).double()
            if ref.norm().item() < 1e-6:
                continue
            got = grads[k[5:]].double().cpu()
            cos = (ref * got).sum().item() / (ref.norm().item() * got.norm().item())
            assert cos > 0.999, (k, cos)


@pytest.mark.parametrize("in_ch,n_cls,nf,size,td,batch", [(3, 2, 32, (96, 128), 4, 2), (1, 4, 16, (48, 48), 8, 1),
                                                        (2, 2, 32, (384, 384), 12, 1)])
def test_fp32_and_bf16_vs_oracle_2d(in_ch, n_cls, nf, size, td, batch):
    """other 2-D configs against the live oracle: nf = 32 (weight-stationary conv, shift-major transposed conv, fused first
    conv on flat volumes), a single modality with 4 classes, and the reference's default 384 x 384 image at td = 12"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = O.param_shapes(in_ch, n_cls, nf, size, td)
    sd = O.synth_state_dict(shapes, seed=3)
    m = HDenseFormer_2D(in_ch, n_cls, nf, size, td)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x = O.synth_mr(batch, in_ch, size, seed=2)
    tgt = O.synth_label(batch, n_cls, size, seed=2)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.forward(sdg, x, td)
    ref_loss = O.deep_super_loss(ref, tgt, ignore_index=0)
    ref_loss.backward()
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    outs = m(x.to(DEV))
    e32 = [rel(o, r.detach()) for o, r in zip(outs, ref)]
    assert max(e32) < 1e-4, e32
    assert (outs[0].argmax(1).cpu() != ref[0].argmax(1)).sum().item() == 0
    loss = crit(outs, tgt.to(DEV))
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    loss.backward()
    worst, wkey = 1.0, None
    for k, p in m.named_parameters():
        r = sdg[k].grad.double()
        if k in ZERO_GRAD_KEYS or r.norm().item() < 1e-7 * max(1.0, sdg[k].double().norm().item()):
            continue
        got = p.grad.double().cpu()
        cosk = (r * got).sum().item() / (r.norm().item() * got.norm().item())
        if cosk < worst:
            worst, wkey = cosk, k
    assert worst > 0.999, (wkey, worst)
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o16 = m(x.to(DEV))
    assert o16[0].dtype == torch.bfloat16
    e16 = rel(o16[0], ref[0].detach())
    # yardstick: the reference graph itself under torch.autocast(bf16) on this GPU (tiny slices leave a few dozen voxels per
    # InstanceNorm statistic at the deep levels, where any bf16 run is noisy)
    sdc = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = O.forward(sdc, x.to(DEV), td)
    ey = rel(y16[0], ref[0].detach())
    assert e16 < max(3e-2, 1.25 * ey), (e16, ey)
    l16 = crit(o16, tgt.to(DEV))
    assert abs(l16.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item())
    l16.backward()
    cs = []
    for k, p in m.named_parameters():
        r = sdg[k].grad.double()
        if k in ZERO_GRAD_KEYS or r.norm().item() < 1e-7 * max(1.0, sdg[k].double().norm().item()):
            continue
        got = p.grad.double().cpu()
        cs.append((r * got).sum().item() / (r.norm().item() * got.norm().item()))
    cs.sort()
    assert cs[len(cs) // 2] > 0.98 and cs[0] > 0.5, (cs[0], cs[len(cs) // 2])      # 0.9885 at the 48 x 48 toy size


def test_2d_train_steps_reduce_loss_and_dropout_runs():
    torch.manual_seed(0)
    size = (64, 64)
    m = HDenseFormer_2D(3, 2, 16, size, 4).to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=2e-3)
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    x = O.synth_mr(2, 3, size, seed=4).to(DEV)
    tgt = O.synth_label(2, 2, size, seed=4).to(DEV)
    losses = []
    for _ in range(8):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = m(x)
        loss = crit(outs, tgt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_2d_graphed_train_step_matches_eager():
    """trainer.GraphedTrainStep is generic: the 2-D module with a capturable torch Adam, dropout off (eval-mode statistics are
    not involved: InstanceNorm has none), replayed steps == eager steps on the same data."""
    from hdenseformer_b200 import trainer as T
    size = (64, 48)
    shapes = O.param_shapes(3, 2, 16, size, 4)
    sd = O.synth_state_dict(shapes, seed=11)
    x = O.synth_mr(2, 3, size, seed=6).to(DEV)
    tgt = O.synth_label(2, 2, size, seed=6).to(DEV)
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    losses = {}
    for mode in ("eager", "graph"):
        m = HDenseFormer_2D(3, 2, 16, size, 4)
        m.load_state_dict(sd)
        m = m.to(DEV).eval()                 # eval: no dropout, so both runs see the same arithmetic
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
        out = []
        if mode == "graph":
            gs = T.GraphedTrainStep(m, crit, opt, x, tgt, use_bf16=False, warmup=1)     # one eager step, then the capture
            out = [float(gs.step(x, tgt).item()) for _ in range(3)]
        else:
            for i in range(4):
                o = m(x)
                loss = crit(o, tgt)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                if i > 0:
                    out.append(float(loss.item()))
        losses[mode] = out
    assert losses["eager"][-1] < losses["eager"][0]
    assert all(abs(a - b) <= 1e-5 * abs(a) for a, b in zip(losses["eager"], losses["graph"])), losses
