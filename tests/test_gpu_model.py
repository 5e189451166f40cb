"""Whole-path parity on the GPU: HDenseFormer forward / loss / backward / sliding window through the C ABI,
against the CPU oracle (oracle/hdf_oracle.py) and the committed golden vectors of the reference.
Tolerances are the north_star's: fp32 logits <= 1e-4 relative, argmax bit-exact, per-tensor gradient cosine
>= 0.999; bf16 logits <= 2e-2."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hdf_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hdenseformer_b200.loss import CEPlusDice, CrossentropyLoss, DeepSuperloss, DiceLoss
    from hdenseformer_b200.models import HDenseFormer
    from hdenseformer_b200 import trainer as T

DEV = "cuda"
ZERO_GRAD_KEYS = ("deep_conv.double_conv.0.bias", "up1.double_conv.0.bias", "up2.double_conv.0.bias",
                  "up3.double_conv.0.bias")   # conv bias before a non-affine InstanceNorm: true gradient is 0


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def build(in_ch, n_cls, nf, size, td, seed=7):
    shapes = O.param_shapes(in_ch, n_cls, nf, size, td)
    sd = O.synth_state_dict(shapes, seed=seed)
    m = HDenseFormer(in_ch, n_cls, nf, image_size=size, transformer_depth=td)
    m.load_state_dict(sd)
    return m.to(DEV), sd


def test_loss_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_cases.npz"))
    for tag in ("c3", "c2", "c4"):
        t = torch.from_numpy(g[f"{tag}_t"]).to(DEV)
        w = torch.from_numpy(g[f"{tag}_w"]).to(DEV)
        for ii, ig in (("ig0", 0), ("ignone", None), ("ig1", 1)):
            for wt, wv in (("w", w), ("nw", None)):
                p = torch.from_numpy(g[f"{tag}_p"]).to(DEV).requires_grad_(True)
                l = CEPlusDice(weight=wv, ignore_index=ig)(p, t)
                l.backward()
                assert abs(l.item() - float(g[f"{tag}_{ii}_{wt}_loss"])) < 2e-6 * max(1, abs(l.item()))
                assert rel(p.grad, torch.from_numpy(g[f"{tag}_{ii}_{wt}_grad"])) < 1e-4
        p = torch.from_numpy(g[f"{tag}_p"]).to(DEV)
        assert abs(DiceLoss(ignore_index=0)(p, t).item() - float(g[f"{tag}_dice_ig0"])) < 2e-6
        assert abs(CrossentropyLoss()(p, t).item() - float(g[f"{tag}_ce"])) < 2e-6
    t = torch.from_numpy(g["ds_t"]).to(DEV)
    outs = [torch.from_numpy(g[f"ds_p{i}"]).to(DEV).requires_grad_(True) for i in range(4)]
    l = DeepSuperloss(CEPlusDice(ignore_index=0))(outs, t)
    l.backward()
    assert abs(l.item() - float(g["ds_loss"])) < 2e-6 * abs(l.item())
    for i, o in enumerate(outs):
        assert rel(o.grad, torch.from_numpy(g[f"ds_g{i}"])) < 1e-4


@pytest.mark.parametrize("name", ["model_nf16_32cube", "model_nf8_aniso"])
def test_fp32_forward_backward_matches_golden(golden_dir, name):
    meta = json.load(open(os.path.join(golden_dir, name + ".json")))
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    size = tuple(meta["image_size"])
    m, _ = build(meta["in_channels"], meta["n_cls"], meta["n_filters"], size, meta["transformer_depth"], meta["param_seed"])
    m.eval()
    M = meta["in_channels"]
    x = O.synth_petct(meta["batch"], size, seed=meta["data_seed"]) if M == 2 else O.synth_mr(meta["batch"], M, size, seed=meta["data_seed"])
    tgt = O.synth_label(meta["batch"], meta["n_cls"], size, seed=meta["data_seed"])
    outs = m(x.to(DEV))
    assert isinstance(outs, list) and len(outs) == 4
    for i, o in enumerate(outs):
        ref = torch.from_numpy(g[f"out{i}"])
        assert tuple(o.shape) == tuple(ref.shape) and o.dtype == torch.float32
        assert rel(o, ref) < 1e-4, (i, rel(o, ref))
    assert torch.equal(outs[0].argmax(1).cpu(), torch.from_numpy(g["out0"]).argmax(1))
    loss = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))(outs, tgt.to(DEV))
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    grads = dict((k, p.grad) for k, p in m.named_parameters())
    for k, (nrm, _) in meta["grad_stats"].items():
        gn = grads[k].double().norm().item()
        if k in ZERO_GRAD_KEYS:
            assert gn < 1e-4
            continue
        assert abs(gn - nrm) <= 2e-3 * nrm + 1e-7, (k, gn, nrm)
    for k in g.files:
        if k.startswith("grad:") and k[5:] not in ZERO_GRAD_KEYS:
            ref = torch.from_numpy(g[k]).double()
            got = grads[k[5:]].double().cpu()
            cos = (ref * got).sum().item() / max(ref.norm().item() * got.norm().item(), 1e-30)
            assert cos > 0.999, (k, cos)


def _oracle_run(sd, x, tgt, td):
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs = O.forward(sdg, x, td)
    loss = O.deep_super_loss(outs, tgt, ignore_index=0)
    loss.backward()
    return [o.detach() for o in outs], loss.item(), {k: v.grad for k, v in sdg.items()}


def test_fp32_and_bf16_vs_oracle_64cube():
    """64^3 (well-conditioned, SURVEY 8c pitfall 2): fp32 gates + bf16 logits; live oracle on CPU."""
    size, td, nf = (64, 64, 64), 4, 16
    m, sd = build(2, 2, nf, size, td)
    m.eval()
    x, tgt = O.synth_petct(1, size, seed=3), O.synth_label(1, 2, size, seed=3)
    ref_outs, ref_loss, ref_g = _oracle_run(sd, x, tgt, td)
    outs = m(x.to(DEV))
    for o, r in zip(outs, ref_outs):
        assert rel(o, r) < 1e-4
    mism = (outs[0].argmax(1).cpu() != ref_outs[0].argmax(1)).sum().item()
    assert mism == 0, f"{mism} argmax mismatches"
    loss = DeepSuperloss(CEPlusDice(ignore_index=0))(outs, tgt.to(DEV))
    assert abs(loss.item() - ref_loss) < 1e-4 * abs(ref_loss)
    loss.backward()
    worst = 1.0
    for k, p in m.named_parameters():
        if k in ZERO_GRAD_KEYS:
            continue
        a, b = p.grad.double().cpu().flatten(), ref_g[k].double().flatten()
        cos = (a @ b).item() / max(a.norm().item() * b.norm().item(), 1e-30)
        worst = min(worst, cos)
        assert cos >= 0.999, (k, cos)
    # bf16 path through autocast like the reference call site (trainer.py:369-370)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs16 = m(x.to(DEV))
    assert outs16[0].dtype == torch.bfloat16
    e = rel(outs16[0], ref_outs[0])
    assert e < 2e-2, e


def test_train_mode_dropout_runs_and_differs():
    size = (32, 32, 32)
    m, _ = build(2, 2, 8, size, 4)
    m.train()
    x = O.synth_petct(1, size, seed=5).to(DEV)
    a = m(x)[0]
    b = m(x)[0]
    assert torch.isfinite(a).all() and not torch.equal(a, b)     # fresh masks each call
    loss = DeepSuperloss(CEPlusDice(ignore_index=0))(m(x), O.synth_label(1, 2, size, seed=5).to(DEV))
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


def test_train_step_reduces_loss():
    size = (32, 32, 32)
    m, _ = build(2, 2, 8, size, 4)
    m.eval()   # deterministic (dropout off) so the loss must go down
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    x, t = O.synth_petct(2, size, seed=9), O.synth_label(2, 2, size, seed=9)
    losses = [T.train_step(m, crit, opt, x, t, use_bf16=False).item() for _ in range(6)]
    assert losses[-1] < losses[0], losses


def test_sliding_window_matches_oracle():
    size, td, nf = (32, 32, 32), 4, 8
    m, sd = build(2, 2, nf, size, td)
    vol = O.synth_petct(1, (48, 40, 32), seed=11)[0]
    ref_mask, ref_prob = O.sliding_window(lambda d: O.forward(sd, d, td)[0], vol, 2, size, (16, 16, 16))
    mask, prob = T.inference_slidingwindow(m, vol, 2, size, (16, 16, 16), use_bf16=False, return_prob=True)
    assert mask.dtype == torch.int64 and tuple(mask.shape) == (48, 40, 32)
    assert rel(prob, ref_prob[0]) < 1e-4
    assert (mask.cpu() != ref_mask).sum().item() == 0
    assert O.mask_dice(mask.cpu(), ref_mask, 2) == 1.0


@pytest.mark.parametrize("in_ch,n_cls,size,batch", [(3, 2, (16, 48, 96), 1), (4, 4, (32, 32, 48), 2)])
def test_bf16_tensor_core_path_other_configs(in_ch, n_cls, size, batch):
    """BASELINE configs 3/4 in miniature (anisotropic 3-modality MR; 4-modality, 4-class, batch 2, deep supervision)
    with nf=16 so every conv takes the tcgen05 path.  These miniature volumes have only 18-24 tokens, so the deep
    InstanceNorms normalise over a handful of voxels and bf16 rounding is amplified (SURVEY 8c pitfall 2: the
    reference's own bf16 autocast is at 1.4e-2 on such sizes); the 2e-2 north-star gate is asserted at 64^3 in
    test_fp32_and_bf16_vs_oracle_64cube and at 96^3 / 144^3 in test_gpu_bench_config.py; here the bound is 4e-2 (the token
    Linears also run on bf16 operands since round 2) plus gradient alignment with the exact fp32 path."""
    td, nf = 4, 16
    m, sd = build(in_ch, n_cls, nf, size, td)
    m.eval()
    x = O.synth_mr(batch, in_ch, size, seed=4)
    tgt = O.synth_label(batch, n_cls, size, seed=4)
    ref = O.forward(sd, x, td)
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = m(x.to(DEV))
    assert outs[0].dtype == torch.bfloat16 and tuple(outs[3].shape) == (batch, n_cls, *(s // 8 for s in size))
    for o, r in zip(outs, ref):
        assert rel(o, r) < 4e-2, rel(o, r)
    crit(outs, tgt.to(DEV)).backward()
    g16 = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    crit(m(x.to(DEV)), tgt.to(DEV)).backward()
    big = [k for k, p in m.named_parameters() if p.numel() >= 4096 and k not in ZERO_GRAD_KEYS]
    cos = []
    for k in big:
        a, b = g16[k].double().flatten(), dict(m.named_parameters())[k].grad.double().flatten()
        cos.append((a @ b).item() / max(a.norm().item() * b.norm().item(), 1e-30))
    # yardstick (SURVEY 8c pitfall 3): the reference's own bf16 autocast reaches min cosine 0.926 vs fp64 on such sizes
    assert min(cos) > 0.90 and sum(cos) / len(cos) > 0.98, (min(cos), sum(cos) / len(cos))


@pytest.mark.parametrize("adamw", [False, True])
def test_fused_adam_matches_torch(adamw):
    """FusedAdam over the gradient arena == torch.optim.Adam/AdamW with the reference's no-decay grouping
    (trainer.py:812-819): 5 steps on identical prescribed gradients (model-produced gradients would feed 1-ulp
    differences back through Adam's normalisation and make the trajectories incomparable), including a learning-rate
    change through param_groups; then one real training step through the trainer API."""
    from hdenseformer_b200.optim import FusedAdam
    size = (32, 32, 32)
    ma, _ = build(2, 2, 8, size, 4)
    mb, _ = build(2, 2, 8, size, 4)
    oa = FusedAdam(ma, lr=1e-3, weight_decay=1e-2, adamw=adamw)
    decay = [p for n, p in mb.named_parameters() if p.dim() > 1 and not n.endswith(".bias")]
    nodec = [p for n, p in mb.named_parameters() if not (p.dim() > 1 and not n.endswith(".bias"))]
    cls = torch.optim.AdamW if adamw else torch.optim.Adam
    ob = cls([{"params": decay, "weight_decay": 1e-2}, {"params": nodec, "weight_decay": 0.0}], lr=1e-3)
    arena = ma._grad_arena()
    gen = torch.Generator(device=DEV).manual_seed(1)
    for it in range(5):
        if it == 2:
            for g in oa.param_groups: g["lr"] = 5e-4
            for g in ob.param_groups: g["lr"] = 5e-4
        for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            g = torch.randn(pa.shape, device=DEV, generator=gen) * (10.0 ** float(torch.randint(-6, 1, (1,)).item()))
            arena.views[k].copy_(g)
            pb.grad = g.clone()
        oa.step()
        ob.step()
    worst, wk = 0.0, None
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        e = ((pa - pb).abs().max() / pb.abs().max().clamp_min(1e-12)).item()
        if e > worst:
            worst, wk = e, k
    assert worst < 5e-6, (wk, worst)
    # and through the trainer API on real gradients: the loss must go down
    ma.eval()
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    x, t = O.synth_petct(2, size, seed=9), O.synth_label(2, 2, size, seed=9)
    losses = [T.train_step(ma, crit, oa, x, t, use_bf16=False).item() for _ in range(5)]
    assert losses[-1] < losses[0], losses
