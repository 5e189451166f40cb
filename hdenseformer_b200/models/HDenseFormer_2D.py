"""Drop-in for the reference's models/HDenseFormer_2D.py: same constructor, same forward(x) -> list of 4
logit tensors [B, C, H/2^i, W/2^i], same state_dict keys / shapes (Conv2d, ConvTranspose2d, 16 x 16 patch embedding).

The 2-D network is the 3-D graph with 2-D layers (models/HDenseFormer_2D.py differs from models/HDenseFormer.py only in
Conv2d / MaxPool2d / bilinear / ConvTranspose2d and the 2-D patch grid), so it runs on the same engine and kernels as flat
volumes [N, 1, H, W, C]:
  * 3 x 3 convolutions = 3 x 3 x 3 convolutions whose only non-zero taps are the kd = 1 plane (the kd = 0 / 2 taps read the
    zero padding of the one-plane volume); the 3-D weight gradient's kd = 1 plane is the 2-D weight gradient;
  * ConvTranspose2d = the even output plane of the 3-D transposed convolution with the same embedding (engine._conv_fwd);
  * MaxPool2d / bilinear x2 / 16 x 16 patches / loss strides: the kernels' depth factor is 1 (hdf_*_ex entries);
  * InstanceNorm, heads, token branch: unchanged (statistics over H x W, rows = H x W voxels).
Parameters keep the reference's 2-D shapes; every step embeds them into 3-D operands (slice copies = host plumbing, no
arithmetic) and slices the gradients back.  The kd = 0 / 2 taps still cost tensor-core time (multiplying zeros): the 2-D
model is a correctness-first widening, not a tuned path.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
from torch import nn

from ..engine import Config, Engine
from .HDenseFormer import _Node, _default_init, param_table


def param_table_2d(in_channels: int, n_cls: int, n_filters: int, image_size, transformer_depth: int) -> Dict[str, tuple]:
    """state_dict key -> shape of the reference 2-D module, in its registration order (models/HDenseFormer_2D.py:172-224)"""
    t3 = param_table(in_channels, n_cls, n_filters, (16, image_size[0], image_size[1]), transformer_depth)
    out: Dict[str, tuple] = {}
    for k, shp in t3.items():
        if len(shp) == 5:
            shp = shp[:2] + shp[3:]                 # drop the depth extent of conv / patch / head kernels
        out[k] = shp
    return out


def _embed(key: str, p2: torch.Tensor) -> torch.Tensor:
    """2-D parameter -> the 3-D operand the engine reads"""
    if p2.dim() != 4:
        return p2
    if p2.shape[2:] == (3, 3):                      # conv / transposed conv: the kd = 1 plane
        w3 = torch.zeros((p2.shape[0], p2.shape[1], 3, 3, 3), dtype=p2.dtype, device=p2.device)
        w3[:, :, 1].copy_(p2)
        return w3
    return p2.unsqueeze(2)                          # 16 x 16 patch kernel, 1 x 1 head: a view with depth extent 1


class _HDF2DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, dtype, training, seed, need_grad, *params):
        P3 = {k: _embed(k, p.detach()) for k, p in zip(module._keys, params)}
        outs, saved = module._engine.forward(P3, x.unsqueeze(2), dtype, training, seed, save=need_grad)
        ctx.module, ctx.saved, ctx.P3 = module, saved, P3
        ctx.set_materialize_grads(False)
        return tuple(o.squeeze(2) for o in outs)

    @staticmethod
    def backward(ctx, *gouts):
        m = ctx.module
        G3 = {k: torch.zeros_like(v) for k, v in ctx.P3.items()}
        gl = [None if g is None else g.unsqueeze(2) for g in gouts]
        m._engine.backward(ctx.P3, G3, ctx.saved, gl, on_grads_ready=None)
        ctx.saved = None
        grads = []
        for k, prm in zip(m._keys, m._param_list()):
            g3 = G3[k]
            if prm.dim() == 4:
                g3 = g3[:, :, 1] if prm.shape[2:] == (3, 3) else g3.squeeze(2)
            grads.append(g3.contiguous() if prm.requires_grad else None)
        return (None,) * 6 + tuple(grads)


class HDenseFormer_2D(nn.Module):
    def __init__(self, in_channels, n_cls, n_filters, image_size=(384, 384), transformer_depth=12):
        super().__init__()
        image_size = tuple(image_size) if isinstance(image_size, (tuple, list)) else (image_size,) * 2
        if len(image_size) != 2 or any(s % 16 != 0 for s in image_size):
            raise ValueError(f"image_size {image_size}: two spatial dims, each a multiple of 16 (patch 16, 4 x2 up-samplings)")
        self.in_channels, self.n_cls, self.n_filters = in_channels, n_cls, n_filters
        self.image_size, self.transformer_depth = image_size, transformer_depth
        self.compute_dtype = None
        table = param_table_2d(in_channels, n_cls, n_filters, image_size, transformer_depth)
        self._keys: List[str] = list(table.keys())
        tensors: Dict[str, torch.Tensor] = {}
        for k, shp in table.items():
            t = _default_init(k, shp)
            if t is None and k.endswith("weight"):
                fan_in = shp[1] * 9 if "upconv_" in k else int(math.prod(shp[1:]))
                bound = 1.0 / math.sqrt(fan_in)
                t = torch.empty(shp).uniform_(-bound, bound)
                bkey = k[:-6] + "bias"
                if bkey in table:
                    tensors[bkey] = torch.empty(table[bkey]).uniform_(-bound, bound)
            if t is not None:
                tensors[k] = t
        for k in self._keys:
            node = self
            parts = k.split(".")
            for name in parts[:-1]:
                if name not in node._modules:
                    node.add_module(name, _Node())
                node = node._modules[name]
            node.register_parameter(parts[-1], nn.Parameter(tensors[k]))
        self._engine = Engine(Config(in_channels, n_cls, n_filters, (1,) + image_size, transformer_depth))
        self._seed_dev = None

    def _param_list(self):
        return [p for _, p in self.named_parameters()]

    def _resolve_dtype(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            if dt == torch.bfloat16:
                return dt
            raise RuntimeError(f"autocast dtype {dt} is not supported by the B200 path (use bfloat16)")
        return torch.float32

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("hdenseformer_b200 has no CPU path: move the model and the input to a B200 (cuda) device")
        x = x.detach().float().contiguous()
        params = self._param_list()
        if self._seed_dev is None or self._seed_dev.device != x.device:
            self._seed_dev = torch.tensor([(torch.initial_seed() * 1000003) & 0x3FFFFFFFFFFFFFFF], dtype=torch.int64,
                                          device=x.device)
        if self.training:
            self._seed_dev.add_(1000003)
        seed = self._seed_dev.clone()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        outs = _HDF2DFunction.apply(self, x, self._resolve_dtype(), self.training, seed, need_grad, *params)
        return list(outs)


def HDenseFormer_2D_32(in_channels, n_cls, image_size, transformer_depth):
    return HDenseFormer_2D(in_channels=in_channels, n_cls=n_cls, image_size=image_size, n_filters=32,
                           transformer_depth=transformer_depth)


def HDenseFormer_2D_16(in_channels, n_cls, image_size, transformer_depth):
    return HDenseFormer_2D(in_channels=in_channels, n_cls=n_cls, image_size=image_size, n_filters=16,
                           transformer_depth=transformer_depth)
