"""Parity on the configurations that are benchmarked: HDenseFormer_32, transformer_depth 12,
BASELINE config 1 (1 x 2 x 96^3) in fp32 and bf16 against the live CPU oracle, config 2 (2 x 2 x 144^3, bf16) forward + loss,
and the step machinery bench.py times: GraphedTrainStep replay == eager train_step, graphed sliding window == eager.

bf16 gradients are reported as the three distances SURVEY 8c asks for (ours-bf16 vs ref-fp32, ref-bf16 vs ref-fp32,
ours-bf16 vs ref-bf16), where ref-bf16 is the reference graph (the oracle's torch ops) run eagerly by cuDNN/cuBLAS under
torch.autocast(bf16) on the same GPU.  The table is printed and written to gpurun_out/r2_parity_table.json."""
import json
import os
import time

import pytest
import torch

from oracle import hdf_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
    from hdenseformer_b200.models import HDenseFormer
    from hdenseformer_b200.optim import FusedAdam
    from hdenseformer_b200 import trainer as T

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZERO_GRAD_KEYS = ("deep_conv.double_conv.0.bias", "up1.double_conv.0.bias", "up2.double_conv.0.bias",
                  "up3.double_conv.0.bias")


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().cpu().abs().max().clamp_min(1e-12)).item()


def build(in_ch, n_cls, nf, size, td, seed=7):
    sd = O.synth_state_dict(O.param_shapes(in_ch, n_cls, nf, size, td), seed=seed)
    m = HDenseFormer(in_ch, n_cls, nf, image_size=size, transformer_depth=td)
    m.load_state_dict(sd)
    return m.to(DEV), sd


def oracle_run(sd, x, tgt, td, device="cpu", autocast=False):
    """reference forward + DeepSuperloss(CEPlusDice(ignore_index=0)) + backward; the loss runs outside autocast on the
    low-precision logits, like trainer.py:369-371"""
    sdg = {k: v.to(device).clone().requires_grad_(True) for k, v in sd.items()}
    x, tgt = x.to(device), tgt.to(device)
    if autocast:
        with torch.autocast(device, dtype=torch.bfloat16):
            outs = O.forward(sdg, x, td)
    else:
        outs = O.forward(sdg, x, td)
    loss = O.deep_super_loss(outs, tgt, ignore_index=0)
    loss.backward()
    return [o.detach() for o in outs], float(loss.item()), {k: v.grad.detach().float().cpu() for k, v in sdg.items()}


def cosines(ga, gb):
    out = {}
    for k in ga:
        if k in ZERO_GRAD_KEYS:
            continue
        a, b = ga[k].double().flatten().cpu(), gb[k].double().flatten().cpu()
        out[k] = (a @ b).item() / max(a.norm().item() * b.norm().item(), 1e-300)
    return out


def summary(cs):
    v = sorted(cs.values())
    return dict(min=v[0], p05=v[len(v) // 20], median=v[len(v) // 2], mean=sum(v) / len(v),
                below_0999=sum(1 for c in v if c < 0.999), n=len(v), argmin=min(cs, key=cs.get))


def dump(name, obj):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "r2_parity_table.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = obj
        json.dump(cur, open(path, "w"), indent=1)
    except OSError:
        pass
    print(f"[parity] {name}: {json.dumps(obj)}")


def test_config1_96cube_nf32_td12_fp32_and_bf16_vs_oracle():
    """BASELINE config 1: HDenseFormer_32(2, 2, 96^3, td=12), batch 1.  fp32 gates of the north star against the CPU
    oracle (logits 1e-4, argmax exact, loss 1e-4, per-tensor gradient cosine >= 0.999); bf16 logits 2e-2 and bf16
    gradients next to the reference-autocast yardstick."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    size, td, nf = (96, 96, 96), 12, 32
    m, sd = build(2, 2, nf, size, td)
    m.eval()
    x, tgt = O.synth_petct(1, size, seed=3), O.synth_label(1, 2, size, seed=3)
    t0 = time.time()
    ref_outs, ref_loss, ref_g = oracle_run(sd, x, tgt, td)
    t_oracle = time.time() - t0
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    # ---- fp32 exact path
    outs = m(x.to(DEV))
    e32 = [rel(o, r) for o, r in zip(outs, ref_outs)]
    mism = (outs[0].argmax(1).cpu() != ref_outs[0].argmax(1)).sum().item()
    loss = crit(outs, tgt.to(DEV))
    loss.backward()
    g32 = {k: p.grad.detach().float().cpu().clone() for k, p in m.named_parameters()}
    c32 = cosines(g32, ref_g)
    zero_norms = {k: g32[k].norm().item() for k in ZERO_GRAD_KEYS}
    m.zero_grad(set_to_none=True)
    # ---- bf16 tensor-core path (autocast like trainer.py:369-370)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs16 = m(x.to(DEV))
    assert outs16[0].dtype == torch.bfloat16
    e16 = [rel(o, r) for o, r in zip(outs16, ref_outs)]
    mism16 = (outs16[0].float().argmax(1).cpu() != ref_outs[0].argmax(1)).sum().item()
    loss16 = crit(outs16, tgt.to(DEV))
    loss16.backward()
    g16 = {k: p.grad.detach().float().cpu().clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    # ---- yardstick: the reference graph, eager cuDNN/cuBLAS, bf16 autocast, same GPU
    y_outs, y_loss, y_g = oracle_run(sd, x, tgt, td, device="cuda", autocast=True)
    ey = [rel(o, r) for o, r in zip(y_outs, ref_outs)]
    tab = dict(
        config="HDenseFormer_32(2,2,96^3,td=12) B=1", oracle_cpu_seconds=round(t_oracle, 2),
        fp32=dict(logit_rel=e32, argmax_mismatch=mism, loss=float(loss.item()), loss_ref=ref_loss, cos=summary(c32)),
        bf16_ours_vs_ref_fp32=dict(logit_rel=e16, argmax_mismatch=mism16, loss=float(loss16.item()), cos=summary(cosines(g16, ref_g))),
        bf16_ref_vs_ref_fp32=dict(logit_rel=ey, loss=y_loss, cos=summary(cosines(y_g, ref_g)),
                                  argmax_mismatch=(y_outs[0].float().argmax(1).cpu() != ref_outs[0].argmax(1)).sum().item()),
        bf16_ours_vs_ref_bf16=dict(logit_rel=[rel(o, r) for o, r in zip(outs16, y_outs)], cos=summary(cosines(g16, y_g))),
    )
    dump("config1_96cube", tab)
    # fp32 gates (north star)
    assert max(e32) < 1e-4, e32
    assert mism == 0, f"{mism} argmax mismatches in fp32"
    assert abs(loss.item() - ref_loss) < 1e-4 * abs(ref_loss)
    assert tab["fp32"]["cos"]["min"] >= 0.999, tab["fp32"]["cos"]
    wn = g32["deep_conv.double_conv.0.weight"].norm().item()
    assert all(v <= 1e-4 * max(wn, 1.0) + 1e-6 for v in zero_norms.values()), zero_norms
    # bf16 gates: logits within the north star's 2e-2; gradients at least as well aligned with the fp32 truth as the
    # reference's own bf16 autocast run (SURVEY 8c pitfall 3: that run itself misses 0.999 on most tensors)
    assert e16[0] < 2e-2, e16
    assert abs(loss16.item() - ref_loss) < 2e-2 * abs(ref_loss)
    ours, yard = tab["bf16_ours_vs_ref_fp32"]["cos"], tab["bf16_ref_vs_ref_fp32"]["cos"]
    assert ours["min"] >= yard["min"] - 0.03, (ours, yard)
    assert ours["mean"] >= yard["mean"] - 0.005, (ours, yard)
    assert ours["median"] >= 0.99, ours


def test_config2_144cube_batch2_bf16_forward_loss_vs_oracle():
    """The benchmarked shape itself (2 x 2 x 144^3, nf=32, td=12, bf16): logits and loss against the CPU fp32 oracle
    (forward only on the CPU: ~10-20 s)."""
    size, td, nf = (144, 144, 144), 12, 32
    m, sd = build(2, 2, nf, size, td)
    m.eval()
    x, tgt = O.synth_petct(2, size, seed=5), O.synth_label(2, 2, size, seed=5)
    t0 = time.time()
    with torch.no_grad():
        ref_outs = O.forward(sd, x, td)
        ref_loss = float(O.deep_super_loss(ref_outs, tgt, ignore_index=0).item())
    t_oracle = time.time() - t0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        outs = m(x.to(DEV))
    loss = DeepSuperloss(CEPlusDice(ignore_index=0))(outs, tgt.to(DEV))
    e = [rel(o, r) for o, r in zip(outs, ref_outs)]
    mism = (outs[0].float().argmax(1).cpu() != ref_outs[0].argmax(1)).sum().item()
    dump("config2_144cube_b2_bf16", dict(logit_rel=e, argmax_mismatch=mism, voxels=int(ref_outs[0].argmax(1).numel()),
                                         loss=float(loss.item()), loss_ref=ref_loss, oracle_cpu_seconds=round(t_oracle, 2)))
    assert e[0] < 2e-2, e
    assert abs(loss.item() - ref_loss) < 2e-2 * abs(ref_loss)
    # random-init logits are near-ties on many voxels: the reference's own bf16 autocast flips 0.43 % of the argmax mask
    # at 64^3 (SURVEY 8c pitfall 3) and 0.31 % at 96^3 (profiles/r2_parity_table.json); ours 0.17 % / 0.56 % here
    assert mism < 2e-2 * ref_outs[0].argmax(1).numel()


@pytest.mark.parametrize("name,in_ch,n_cls,size", [
    ("config3_3mod_32x384x384", 3, 2, (32, 384, 384)),      # PI-CAI-shaped MR volume (depth 24 zero-padded to 32)
    ("config4_4mod_128cube", 4, 4, (128, 128, 128)),        # BraTS-shaped: 4 modalities, 4 classes, deep supervision
])
def test_mr_configs_bf16_forward_loss_backward_vs_oracle(name, in_ch, n_cls, size):
    """BASELINE configs 3 and 4 at their full sizes (nf=32, td=12, one sample): bf16 logits, loss and per-tensor gradient
    cosines against the fp32 oracle, which runs on the same GPU here (cuDNN/cuBLAS fp32 with TF32 off) because a CPU
    backward at these sizes takes minutes.  512->256 / 4-modality stems and 4-class heads and losses are only reached here."""
    td, nf = 12, 32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = build(in_ch, n_cls, nf, size, td)
    m.eval()
    x, tgt = O.synth_mr(1, in_ch, size, seed=9), O.synth_label(1, n_cls, size, seed=9)
    ref_outs, ref_loss, ref_g = oracle_run(sd, x, tgt, td, device=DEV)
    torch.cuda.empty_cache()
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = m(x.to(DEV))
    loss = DeepSuperloss(CEPlusDice(ignore_index=0))(outs, tgt.to(DEV))
    loss.backward()
    g = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()}
    loss_v = float(loss.item())
    e = [rel(o, r) for o, r in zip(outs, ref_outs)]
    cs = summary(cosines(g, ref_g))
    mism = (outs[0].float().argmax(1) != ref_outs[0].argmax(1)).sum().item()
    del outs, loss
    m.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    # yardstick: the reference graph under torch.autocast(bf16) on the same GPU (SURVEY 8c three-distance table)
    y_outs, y_loss, y_g = oracle_run(sd, x, tgt, td, device=DEV, autocast=True)
    ey = [rel(o, r) for o, r in zip(y_outs, ref_outs)]
    ycs = summary(cosines(y_g, ref_g))
    ymism = (y_outs[0].float().argmax(1) != ref_outs[0].argmax(1)).sum().item()
    dump(name, dict(voxels=int(ref_outs[0][:, 0].numel()), loss_ref=ref_loss,
                    bf16_ours_vs_ref_fp32=dict(logit_rel=e, argmax_mismatch=mism, loss=loss_v, grad_cos=cs),
                    bf16_ref_vs_ref_fp32=dict(logit_rel=ey, argmax_mismatch=ymism, loss=y_loss, grad_cos=ycs)))
    # MR-shaped inputs (non-negative, 3-4 channels) put the max-norm logit error of ANY bf16 run above the 2e-2 the PET/CT
    # configs meet: the gate is "no worse than the reference's own autocast run" (its numbers are stored next to ours)
    assert e[0] < max(2e-2, 1.25 * ey[0]), (e, ey)
    assert abs(loss_v - ref_loss) < 2e-2 * abs(ref_loss)
    assert mism < max(2e-2 * ref_outs[0][:, 0].numel(), 1.25 * ymism), (mism, ymism)
    assert cs["median"] >= min(0.99, ycs["median"] - 0.002) and cs["mean"] >= ycs["mean"] - 0.005, (cs, ycs)
    assert cs["min"] >= ycs["min"] - 0.05, (cs, ycs)


def _reset(m, sd, opt, opt_sd):
    with torch.no_grad():
        for k, p in m.named_parameters():
            p.copy_(sd[k].to(p.device))
    opt.load_state_dict(opt_sd)


@pytest.mark.parametrize("use_bf16", [False, True])
def test_graphed_train_step_equals_eager(use_bf16):
    """GraphedTrainStep (what bench.py times) replays exactly what eager train_step computes: 3 steps with FusedAdam,
    dropout off, different batches, plain and prefetched H2D paths, including a learning-rate change between replays."""
    size, td, nf = (32, 32, 32), 4, 16
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    batches = [(O.synth_petct(2, size, seed=30 + i).pin_memory(), O.synth_label(2, 2, size, seed=30 + i).pin_memory())
               for i in range(3)]
    lrs = [1e-3, 1e-3, 2.5e-4]
    # eager
    ma, sd = build(2, 2, nf, size, td)
    ma.eval()
    oa = FusedAdam(ma, lr=1e-3, weight_decay=1e-4)
    la = []
    for (x, t), lr in zip(batches, lrs):
        for g in oa.param_groups:
            g["lr"] = lr
        la.append(T.train_step(ma, crit, oa, x, t, use_bf16=use_bf16).item())
    # graphed: the constructor warms up with real steps, so parameters and optimizer state are restored afterwards
    mb, _ = build(2, 2, nf, size, td)
    mb.eval()
    ob = FusedAdam(mb, lr=1e-3, weight_decay=1e-4)
    osd = ob.state_dict()
    gs = T.GraphedTrainStep(mb, crit, ob, batches[0][0], batches[0][1], use_bf16=use_bf16)
    _reset(mb, sd, ob, osd)
    lb = []
    for i, ((x, t), lr) in enumerate(zip(batches, lrs)):
        for g in ob.param_groups:
            g["lr"] = lr                      # what an LR scheduler does; step() must push it to the device
        if i == 1:
            gs.prefetch(x, t)                 # prefetched H2D path
        lb.append(gs.step(x, t).item())
    torch.cuda.synchronize()
    assert all(abs(a - b) <= 1e-6 * abs(a) for a, b in zip(la, lb)), (la, lb)
    worst = max(((pa - pb).abs().max() / pa.abs().max().clamp_min(1e-12)).item()
                for pa, pb in zip(ma.parameters(), mb.parameters()))
    assert worst <= 1e-6, worst
    assert la[-1] < la[0]


def test_lr_scheduler_drives_fused_adam():
    """FusedAdam must be a torch.optim.Optimizer so that torch LR schedulers accept it (trainer.py:263-264)."""
    size = (32, 32, 32)
    m, _ = build(2, 2, 8, size, 4)
    m.eval()
    opt = FusedAdam(m, lr=1e-2)
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 2], gamma=0.1)
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    x, t = O.synth_petct(1, size, seed=1), O.synth_label(1, 2, size, seed=1)
    seen = []
    for _ in range(3):
        T.train_step(m, crit, opt, x, t, use_bf16=False)
        seen.append(float(opt.hyper[0].item()))
        sched.step()
    assert seen == pytest.approx([1e-2, 1e-3, 1e-4], rel=1e-6), seen


def test_gradient_accumulation_and_double_forward():
    """a second backward without zero_grad accumulates (it used to wipe the first gradients), and a model
    called twice inside one autograd graph receives the sum of both gradients."""
    size = (32, 32, 32)
    m, _ = build(2, 2, 8, size, 4)
    m.eval()
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    xa, ta = O.synth_petct(1, size, seed=1).to(DEV), O.synth_label(1, 2, size, seed=1).to(DEV)
    xb, tb = O.synth_petct(1, size, seed=2).to(DEV), O.synth_label(1, 2, size, seed=2).to(DEV)

    def grads():
        return {k: p.grad.detach().clone() for k, p in m.named_parameters()}

    crit(m(xa), ta).backward()
    ga = grads()
    m.zero_grad(set_to_none=True)
    crit(m(xb), tb).backward()
    gb = grads()
    m.zero_grad(set_to_none=True)
    # two backwards, no zero_grad in between
    crit(m(xa), ta).backward()
    crit(m(xb), tb).backward()
    g2 = grads()
    m.zero_grad(set_to_none=True)
    # two forwards, one backward
    (crit(m(xa), ta) + crit(m(xb), tb)).backward()
    g3 = grads()
    for k in ga:
        ref = ga[k] + gb[k]
        den = ref.abs().max().clamp_min(1e-12)
        assert ((g2[k] - ref).abs().max() / den).item() < 1e-5, k
        assert ((g3[k] - ref).abs().max() / den).item() < 1e-5, k


def test_graphed_sliding_window_equals_eager():
    size, td, nf = (32, 32, 32), 4, 16
    m, sd = build(2, 2, nf, size, td)
    vol = O.synth_petct(1, (48, 40, 32), seed=11)[0]
    for bf in (False, True):
        ma, pa = T.inference_slidingwindow(m, vol, 2, size, (16, 16, 16), use_bf16=bf, return_prob=True, use_graph=False)
        mb, pb = T.inference_slidingwindow(m, vol, 2, size, (16, 16, 16), use_bf16=bf, return_prob=True, use_graph=True)
        assert torch.equal(ma, mb)
        assert rel(pb, pa) <= 1e-6
        mc, pc = T.inference_slidingwindow(m, vol, 2, size, (16, 16, 16), use_bf16=bf, return_prob=True, use_graph=True,
                                           patch_batch=2)
        assert torch.equal(ma, mc)
        assert rel(pc, pa) <= 1e-6
    ref_mask, _ = O.sliding_window(lambda d: O.forward(sd, d, td)[0], vol, 2, size, (16, 16, 16))
    assert (mb.cpu() != ref_mask).float().mean().item() < 2e-2      # bf16 mask vs the fp32 oracle mask
