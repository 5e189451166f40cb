"""autograd bridge for the fused Dice + cross-entropy kernels (csrc/loss.cu)."""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import ops


_CW_CACHE = {}


def _device_weight(cw: Optional[torch.Tensor], dev) -> Optional[torch.Tensor]:
    """fp32 device copy of the class weights, made once per (tensor, version, device): the reference keeps the weight
    as a plain CPU tensor attribute (trainer.py:743-744; not a registered buffer), and a fresh pageable-host -> device
    copy on every step would synchronise the stream and cannot be captured in a CUDA graph."""
    if cw is None:
        return None
    cw = cw.detach()
    if cw.device == dev and cw.dtype == torch.float32 and cw.is_contiguous():
        return cw
    key = (cw.data_ptr(), cw._version, tuple(cw.shape), str(cw.dtype), str(dev))
    hit = _CW_CACHE.get(key)
    if hit is None:
        if len(_CW_CACHE) > 16:
            _CW_CACHE.clear()
        hit = (cw, cw.to(device=dev, dtype=torch.float32).contiguous())     # keep the source alive: data_ptr stays unique
        _CW_CACHE[key] = hit
    return hit[1]


class _SegLossFn(torch.autograd.Function):
    """loss = sum_i level_weight_i * (ce_w * CE(pred_i, nearest(target)) + dice_w * Dice(pred_i, nearest(target)))."""

    @staticmethod
    def forward(ctx, target, cw, ignore_index, smooth, ce_w, dice_w, level_weights, levels, *preds):
        ops.ensure_init(preds[0])
        dev = preds[0].device
        B, C = preds[0].shape[:2]
        nbytes = ops._lib().hdf_loss_sums_bytes(B, C)
        total = torch.zeros(1, dtype=torch.float32, device=dev)
        per_level = torch.zeros((len(preds), 3), dtype=torch.float32, device=dev)
        sums: List[torch.Tensor] = []
        target = target.detach().float().contiguous()
        cwd = _device_weight(cw, dev)
        ps = []
        for i, p in enumerate(preds):
            p = p.detach()
            if p.dtype not in (torch.float32, torch.bfloat16):
                p = p.float()
            p = p.contiguous()
            s = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
            ops.loss_level_fwd(p, target, cwd, levels[i], ignore_index, smooth, level_weights[i], ce_w, dice_w, s,
                               per_level[i], total)
            sums.append(s)
            ps.append(p)
        ctx.save_for_backward(target, *ps, *sums)
        ctx.meta = (cwd, ignore_index, smooth, ce_w, dice_w, level_weights, levels, [p.dtype for p in preds])
        ctx.per_level = per_level
        return total[0]

    @staticmethod
    def backward(ctx, gout):
        cwd, ignore_index, smooth, ce_w, dice_w, level_weights, levels, in_dtypes = ctx.meta
        saved = ctx.saved_tensors
        n = len(levels)
        target, ps, sums = saved[0], saved[1:1 + n], saved[1 + n:]
        gout = gout.detach().float().reshape(1).contiguous()
        grads = []
        for i in range(n):
            if not ctx.needs_input_grad[8 + i]:
                grads.append(None)
                continue
            d = torch.empty_like(ps[i])
            ops.loss_level_bwd(ps[i], target, cwd, levels[i], ignore_index, smooth, level_weights[i], ce_w, dice_w, sums[i],
                               gout, d)
            grads.append(d if d.dtype == in_dtypes[i] else d.to(in_dtypes[i]))
        return (None,) * 8 + tuple(grads)


def seg_loss(preds, target, weight=None, ignore_index=None, smooth=1e-5, ce_w=1.0, dice_w=1.0, level_weights=None,
             levels=None):
    if not preds[0].is_cuda:
        raise RuntimeError("hdenseformer_b200 losses have no CPU path: tensors must be on a B200 (cuda) device")
    n = len(preds)
    level_weights = level_weights or [1.0] * n
    levels = levels or [0] * n
    if target.dim() == 4:       # 2-D model (models/HDenseFormer_2D.py): [B, C, H, W] -> flat volumes [B, C, 1, H, W]
        target = target.unsqueeze(2)
        preds = [p.unsqueeze(2) for p in preds]
    D, H, W = target.shape[2:]
    for p, lv in zip(preds, levels):
        exp = (target.shape[0], target.shape[1], D if D == 1 else D >> lv, H >> lv, W >> lv)
        assert tuple(p.shape) == exp, f"predict {tuple(p.shape)} vs target level {lv} {exp}"
    return _SegLossFn.apply(target, weight, ignore_index, float(smooth), float(ce_w), float(dice_w), list(level_weights),
                            list(levels), *preds)
