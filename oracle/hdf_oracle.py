"""CPU oracle for the H-DenseFormer 3D hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch ops, any float dtype, no
nn.Module, no kernels of ours) of the reference algorithm.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it; the product package `hdenseformer_b200` never does.

Pinning: the reference ships no tests/golden vectors, so this
restatement is pinned against outputs of the *unmodified reference modules*
imported from /root/reference in the build container.  The generating script is
`tests/golden/make_golden.py`; the vectors it wrote are `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against them on every CPU run.

Every function cites the reference lines it follows (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# deterministic synthetic parameters / inputs (shared by tests, bench, golden)
# --------------------------------------------------------------------------
def param_shapes(in_channels: int, n_cls: int, n_filters: int, image_size: Sequence[int],
                 transformer_depth: int) -> Dict[str, tuple]:
    """Key -> shape table of the reference state_dict (models/HDenseFormer.py:178-227,
    :104-130, :78-89).  Ordered like the reference's registration order."""
    nf = n_filters
    E = 4 * nf
    g = 32  # growth_rate default, models/HDenseFormer.py:79
    nd = len(image_size)          # 3: models/HDenseFormer.py; 2: models/HDenseFormer_2D.py (same graph, 2-D layers)
    N = int(np.prod([s // 16 for s in image_size]))
    k3, k16, k1 = (3,) * nd, (16,) * nd, (1,) * nd
    out: Dict[str, tuple] = {}
    for i in range(in_channels):
        p = f"attns.{i}."
        out[p + "position_embeddings"] = (1, N, E)
        out[p + "patch_embeddings.weight"] = (E, 1) + k16
        out[p + "patch_embeddings.bias"] = (E,)
        for b in range(transformer_depth // 4):
            q = p + f"blocks.{b}.0."
            for l in range(4):
                r = q + f"layers.{l}."
                out[r + "0.weight"] = (g, E + l * g)
                out[r + "0.bias"] = (g,)
                out[r + "1.norm.weight"] = (g,)
                out[r + "1.norm.bias"] = (g,)
                out[r + "1.fn.to_qkv.weight"] = (3 * g, g)
                out[r + "1.fn.to_out.0.weight"] = (g, g)
                out[r + "1.fn.to_out.0.bias"] = (g,)
                out[r + "2.norm.weight"] = (g,)
                out[r + "2.norm.bias"] = (g,)
                out[r + "2.fn.net.0.weight"] = (2 * g, g)
                out[r + "2.fn.net.0.bias"] = (2 * g,)
                out[r + "2.fn.net.3.weight"] = (g, 2 * g)
                out[r + "2.fn.net.3.bias"] = (g,)
            out[q + "out_layer.net.0.weight"] = (2 * g, E + 4 * g)
            out[q + "out_layer.net.0.bias"] = (2 * g,)
            out[q + "out_layer.net.3.weight"] = (E, 2 * g)
            out[q + "out_layer.net.3.bias"] = (E,)

    def upconv(name, ci, co):
        out[f"{name}.double_conv.0.weight"] = (co, ci) + k3
        out[f"{name}.double_conv.0.bias"] = (co,)

    def basic(name, ci, co):
        out[f"{name}.conv.weight"] = (co, ci) + k3
        out[f"{name}.norm.weight"] = (co,)
        out[f"{name}.norm.bias"] = (co,)

    def convt(name, ci, co):
        out[f"{name}.weight"] = (ci, co) + k3
        out[f"{name}.bias"] = (co,)

    def head(name, ci):
        out[f"{name}.weight"] = (n_cls, ci) + k1
        out[f"{name}.bias"] = (n_cls,)

    upconv("deep_conv", E * in_channels, 8 * nf)
    upconv("up1", 8 * nf, 4 * nf)
    upconv("up2", 4 * nf, 2 * nf)
    upconv("up3", 2 * nf, nf)
    basic("block_1_1_left", in_channels, nf)
    basic("block_1_2_left", nf, nf)
    basic("block_2_1_left", nf, 2 * nf)
    basic("block_2_2_left", 2 * nf, 2 * nf)
    basic("block_3_1_left", 2 * nf, 4 * nf)
    basic("block_3_2_left", 4 * nf, 4 * nf)
    basic("block_4_1_left", 4 * nf, 8 * nf)
    basic("block_4_2_left", 8 * nf, 8 * nf)
    convt("upconv_3", 8 * nf, 4 * nf)
    basic("block_3_1_right", 8 * nf, 4 * nf)
    basic("block_3_2_right", 4 * nf, 4 * nf)
    convt("upconv_2", 4 * nf, 2 * nf)
    basic("block_2_1_right", 4 * nf, 2 * nf)
    basic("block_2_2_right", 2 * nf, 2 * nf)
    convt("upconv_1", 2 * nf, nf)
    basic("block_1_1_right", 2 * nf, nf)
    basic("block_1_2_right", nf, nf)
    head("conv1x1", nf)
    head("conv1x1_d1", 2 * nf)
    head("conv1x1_d2", 4 * nf)
    head("conv1x1_d3", 8 * nf)
    return out


def synth_state_dict(shapes: Dict[str, tuple], seed: int = 0, dtype=torch.float32) -> StateDict:
    """Deterministic parameters that exercise every term (position embeddings and
    biases non-zero, norm weights != 1; SURVEY.md 8c pitfall 4).  Fan-in scaled so
    activations stay O(1) through the depth of the net."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for k, shp in shapes.items():
        if k.endswith("position_embeddings"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif k.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("norm.bias"):
            t = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = int(np.prod(shp[1:]))
            if "upconv_" in k:  # ConvTranspose weight is [Cin,Cout,k,..]; each output sees <= 2^nd taps
                fan_in = shp[0] * 2 ** (len(shp) - 2)
            t = torch.randn(shp, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[k] = t.to(dtype)
    return sd


def synth_petct(batch: int, size: Sequence[int], seed: int = 0) -> torch.Tensor:
    """PET/CT-shaped input [B,2,D,H,W] f32 (value ranges of data_utils/data_loader.py:53-68;
    SURVEY.md 8d): ch0 'CT' = clip(N(0,.35),-1,1); ch1 'PET' = z-scored exp(N(0,1))."""
    g = torch.Generator().manual_seed(1000 + seed)
    ct = (0.35 * torch.randn((batch, 1, *size), generator=g)).clamp_(-1, 1)
    pet = torch.exp(torch.randn((batch, 1, *size), generator=g))
    pet = (pet - pet.mean()) / (pet.std() + 1e-3)
    return torch.cat([ct, pet], 1).float()


def synth_mr(batch: int, channels: int, size: Sequence[int], seed: int = 0) -> torch.Tensor:
    """MR-shaped input [B,M,D,H,W] f32: non-negative, max 1 per channel
    (data_utils/data_loader.py:39-50)."""
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.rand((batch, channels, *size), generator=g) ** 2
    return (x / x.amax(dim=tuple(range(2, 2 + len(size))), keepdim=True)).float()


def synth_label(batch: int, n_cls: int, size: Sequence[int], seed: int = 0) -> torch.Tensor:
    """One-hot f32 label [B,C,D,H,W], channel 0 = background, one seeded ellipsoid per
    foreground class (contract of data_utils/data_loader.py:146-151)."""
    rng = np.random.RandomState(3000 + seed)
    grids = np.meshgrid(*[np.arange(s) for s in size], indexing="ij")      # 3-D volumes or 2-D slices
    lab = np.zeros((batch, *size), dtype=np.int64)
    for b in range(batch):
        for c in range(1, n_cls):
            ctr = [rng.uniform(0.3, 0.7) * s for s in size]
            rad = [rng.uniform(0.12, 0.25) * s for s in size]
            m = sum(((g - ctr[i]) / rad[i]) ** 2 for i, g in enumerate(grids)) <= 1
            lab[b][m] = c
    oh = np.stack([(lab == c) for c in range(n_cls)], 1).astype(np.float32)
    return torch.from_numpy(oh)


# --------------------------------------------------------------------------
# model forward  (models/HDenseFormer.py)
# --------------------------------------------------------------------------
def _dense_forward(sd, p, x, drop):
    """DenseForward: Linear -> exact GELU -> Dropout -> Linear -> Dropout
    (models/HDenseFormer.py:33-44)."""
    x = F.linear(x, sd[p + "net.0.weight"], sd[p + "net.0.bias"])
    x = drop(F.gelu(x))
    x = F.linear(x, sd[p + "net.3.weight"], sd[p + "net.3.bias"])
    return drop(x)


def _dense_attention(sd, p, x, drop, heads=8):
    """Dense_Attention: qkv (no bias), 8 heads of dim 4, softmax(q k^T * d^-.5) v,
    to_out Linear + Dropout (models/HDenseFormer.py:47-75)."""
    B, N, C = x.shape
    dh = C // heads
    qkv = F.linear(x, sd[p + "to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (t.reshape(B, N, heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    dots = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)
    attn = dots.softmax(dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, N, C)
    return drop(F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"]))


def _dct_block(sd, p, x, drop):
    """DensePreConv_AttentionBlock.forward (models/HDenseFormer.py:91-101).  The ff
    module is applied twice with shared weights (:97-98)."""
    feats = [x]
    for l in range(4):
        q = p + f"layers.{l}."
        x = torch.cat(feats, 2)
        x = F.linear(x, sd[q + "0.weight"], sd[q + "0.bias"])
        ln = lambda t, r: F.layer_norm(t, (t.shape[-1],), sd[q + r + ".norm.weight"], sd[q + r + ".norm.bias"], 1e-5)
        x = _dense_attention(sd, q + "1.fn.", ln(x, "1"), drop) + x
        x = _dense_forward(sd, q + "2.fn.", ln(x, "2"), drop) + x
        feats.append(_dense_forward(sd, q + "2.fn.", ln(x, "2"), drop))
    x = torch.cat(feats, 2)
    return _dense_forward(sd, p + "out_layer.", x, drop)


def _transformer_branch(sd, p, img, n_blocks, drop):
    """Dense_TransformerBlock.forward (models/HDenseFormer.py:132-145).  The trailing
    F.interpolate to the same size is nearest -> identity (SURVEY.md 2.1 K8')."""
    x = _conv(img, sd[p + "patch_embeddings.weight"], sd[p + "patch_embeddings.bias"], stride=16)
    shp = x.shape
    x = x.flatten(2).transpose(-1, -2)
    x = drop(x + sd[p + "position_embeddings"])
    for b in range(n_blocks):
        x = _dct_block(sd, p + f"blocks.{b}.0.", x, drop)
    return x.transpose(1, 2).reshape(shp)


def _conv(x, w, b=None, **kw):
    """nn.Conv3d (models/HDenseFormer.py) or nn.Conv2d (models/HDenseFormer_2D.py, the same graph with 2-D layers)"""
    return (F.conv2d if x.dim() == 4 else F.conv3d)(x, w, b, **kw)


def _pool(x):
    return F.max_pool2d(x, 2, 2) if x.dim() == 4 else F.max_pool3d(x, 2, 2)


def _basic(sd, name, x):
    """BasicConv3d: conv k3 p1 no bias -> InstanceNorm3d(affine) -> ReLU
    (models/HDenseFormer.py:148-159)."""
    x = _conv(x, sd[name + ".conv.weight"], None, padding=1)
    x = F.instance_norm(x, weight=sd[name + ".norm.weight"], bias=sd[name + ".norm.bias"], eps=1e-5)
    return F.relu(x)


def _upconv(sd, name, x):
    """UpConv: conv k3 p1 bias -> InstanceNorm3d(no affine) -> ReLU -> trilinear x2
    (models/HDenseFormer.py:162-175)."""
    x = _conv(x, sd[name + ".double_conv.0.weight"], sd[name + ".double_conv.0.bias"], padding=1)
    x = F.relu(F.instance_norm(x, eps=1e-5))
    return F.interpolate(x, scale_factor=2, mode="bilinear" if x.dim() == 4 else "trilinear", align_corners=False)


def _convt(sd, name, x):
    """nn.ConvTranspose3d k3 s2 p1 op1 (models/HDenseFormer.py:211,215,219)."""
    ct = F.conv_transpose2d if x.dim() == 4 else F.conv_transpose3d
    return ct(x, sd[name + ".weight"], sd[name + ".bias"], stride=2, padding=1, output_padding=1)


def _head(sd, name, x):
    return _conv(x, sd[name + ".weight"], sd[name + ".bias"])


def forward(sd: StateDict, x: torch.Tensor, transformer_depth: int = 12, dropout_p: float = 0.0,
            generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
    """HDenseFormer.forward (models/HDenseFormer.py:229-255).  dropout_p=0 is the
    eval()/parity mode (eval differs from train only by dropout)."""
    M = x.shape[1]
    if dropout_p > 0:
        def drop(t):
            keep = (torch.rand(t.shape, generator=generator, dtype=torch.float32) >= dropout_p).to(t.dtype)
            return t * keep / (1 - dropout_p)
    else:
        drop = lambda t: t
    nb = transformer_depth // 4
    attnall = torch.cat([_transformer_branch(sd, f"attns.{i}.", x[:, i:i + 1], nb, drop) for i in range(M)], 1)
    attnout = _upconv(sd, "deep_conv", attnall)
    at1 = _upconv(sd, "up1", attnout)
    at2 = _upconv(sd, "up2", at1)
    at3 = _upconv(sd, "up3", at2)
    ds0 = _basic(sd, "block_1_2_left", _basic(sd, "block_1_1_left", x)) + at3
    ds1 = _basic(sd, "block_2_2_left", _basic(sd, "block_2_1_left", _pool(ds0))) + at2
    ds2 = _basic(sd, "block_3_2_left", _basic(sd, "block_3_1_left", _pool(ds1))) + at1
    y = _basic(sd, "block_4_2_left", _basic(sd, "block_4_1_left", _pool(ds2))) + attnout
    out3 = _head(sd, "conv1x1_d3", y)
    y = _basic(sd, "block_3_2_right", _basic(sd, "block_3_1_right", torch.cat([_convt(sd, "upconv_3", y), ds2], 1)))
    out2 = _head(sd, "conv1x1_d2", y)
    y = _basic(sd, "block_2_2_right", _basic(sd, "block_2_1_right", torch.cat([_convt(sd, "upconv_2", y), ds1], 1)))
    out1 = _head(sd, "conv1x1_d1", y)
    y = _basic(sd, "block_1_2_right", _basic(sd, "block_1_1_right", torch.cat([_convt(sd, "upconv_1", y), ds0], 1)))
    return [_head(sd, "conv1x1", y), out1, out2, out3]


# --------------------------------------------------------------------------
# loss  (loss/combine_loss.py, loss/dice_loss.py, loss/cross_entropy.py)
# --------------------------------------------------------------------------
def dice_loss(predict, target, weight=None, ignore_index=None, smooth=1e-5, p=1):
    """DiceLoss + BinaryDiceLoss(reduction='mean') (loss/dice_loss.py:70-87, :26-41)."""
    C = target.shape[1]
    prob = F.softmax(predict, dim=1)
    total = 0
    for i in range(C):
        if i != ignore_index:
            pi = prob[:, i].reshape(prob.shape[0], -1)
            ti = target[:, i].reshape(target.shape[0], -1)
            inter = (pi * ti).sum(1)
            union = (pi.pow(p) + ti.pow(p)).sum(1)
            l = (1 - (2 * inter + smooth) / (union + smooth)).mean()
            if weight is not None:
                l = l * weight[i]
            total = total + l
    return total / (C - 1) if ignore_index is not None else total / C


def ce_loss(predict, target, weight=None):
    """CrossentropyLoss.forward (loss/cross_entropy.py:10-22): argmax of the one-hot,
    voxel-mean nn.CrossEntropyLoss."""
    tgt = torch.argmax(target, 1) if target.shape[1] > 1 else target[:, 0]
    C = predict.shape[1]
    inp = predict.movedim(1, -1).reshape(-1, C)
    return F.cross_entropy(inp, tgt.long().reshape(-1), weight=weight)


def ce_plus_dice(predict, target, weight=None, ignore_index=None, **kw):
    """CEPlusDice.forward (loss/combine_loss.py:25-35)."""
    assert predict.size() == target.size()
    return ce_loss(predict, target, weight) + dice_loss(predict, target, weight, ignore_index, **kw)


def deep_super_loss(outputs, target, weight=None, ignore_index=0, **kw):
    """DeepSuperloss.forward (loss/combine_loss.py:72-79): sum_i 2^-i * criterion(out_i,
    nearest-resized one-hot target)."""
    loss = 0
    for i, img in enumerate(outputs):
        label = F.interpolate(target, img.shape[2:])
        loss = loss + ce_plus_dice(img, label, weight, ignore_index, **kw) * (1 / (2 ** i))
    return loss


# --------------------------------------------------------------------------
# sliding window  (trainer.py:488-618) and metric (trainer.py:891-945)
# --------------------------------------------------------------------------
def cal_steps(image_size, patch_size, step_size):
    """SemanticSeg.cal_steps (trainer.py:595-618)."""
    steps = []
    for dim in range(len(image_size)):
        if image_size[dim] <= patch_size[dim]:
            steps.append([0])
        else:
            mx = image_size[dim] - patch_size[dim]
            n = int(np.ceil(mx / step_size[dim])) + 1
            act = mx / (n - 1)
            steps.append([int(np.round(act * i)) for i in range(n)])
    return steps


def sliding_window(net_fn, image: torch.Tensor, n_cls: int, patch_size, step_size):
    """inference_slidingwindow hot loop (trainer.py:521-582).  `net_fn(patch[1,M,*])`
    returns full-resolution logits [1,C,*].  Returns (argmax mask [X,Y,Z] int64,
    averaged probabilities [1,C,X,Y,Z])."""
    M, X, Y, Z = image.shape
    agg = torch.zeros((1, n_cls, X, Y, Z), dtype=torch.float32)
    cnt = torch.zeros((1, n_cls, X, Y, Z), dtype=torch.float32)
    steps = cal_steps((X, Y, Z), patch_size, step_size)
    for x in steps[0]:
        ux = min(x + patch_size[0], X)
        for y in steps[1]:
            uy = min(y + patch_size[1], Y)
            for z in steps[2]:
                uz = min(z + patch_size[2], Z)
                data = image[None, :, x:ux, y:uy, z:uz]
                prob = F.softmax(net_fn(data).float(), dim=1)
                prob = F.interpolate(prob, (ux - x, uy - y, uz - z))
                agg[:, :, x:ux, y:uy, z:uz] += prob
                cnt[:, :, x:ux, y:uy, z:uz] += 1
    out = agg / cnt
    return torch.argmax(torch.softmax(out, dim=1), 1)[0], out


def compute_dice(predict, target, ignore_index=0, smooth=1e-5):
    """compute_dice / binary_dice (trainer.py:891-945): hard-mask Dice, mean over
    classes 1.. (classes absent from both masks count as 1)."""
    op = torch.argmax(F.softmax(predict, dim=1), dim=1)
    ot = torch.argmax(target, dim=1)
    C = target.shape[1]
    dl = np.ones((C,), dtype=np.float32)
    for i in range(C):
        if i != ignore_index:
            if not (op == i).any() and not (ot == i).any():
                continue
            a = (op == i).float().reshape(op.shape[0], -1)
            b = (ot == i).float().reshape(ot.shape[0], -1)
            d = ((2 * (a * b).sum(1) + smooth) / ((a + b).sum(1) + smooth)).mean()
            dl[i] = round(d.item(), 4)
    return float(np.nanmean(dl[1:]))


def mask_dice(a: torch.Tensor, b: torch.Tensor, n_cls: int) -> float:
    """Dice between two integer masks, mean over foreground classes present in either."""
    ds = []
    for c in range(1, n_cls):
        x, y = (a == c), (b == c)
        den = x.sum().item() + y.sum().item()
        if den == 0:
            continue
        ds.append(2.0 * (x & y).sum().item() / den)
    return float(np.mean(ds)) if ds else 1.0
