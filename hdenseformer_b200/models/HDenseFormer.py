"""Drop-in for the reference's models/HDenseFormer.py: same constructor, same forward(x) -> list of 4
logit tensors, same state_dict keys/shapes (models/HDenseFormer.py:178-261).

The module holds parameters only; all arithmetic runs in libhdf_b200 kernels through
hdenseformer_b200.engine.Engine.  Compute dtype follows the caller like the reference does: inside
`torch.autocast('cuda', torch.bfloat16)` activations are bf16 (and the returned logits are bf16, as the
reference's would be), otherwise everything is fp32 with no TF32 anywhere.  Set `model.compute_dtype`
to force one.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
from torch import nn

from ..engine import Config, Engine, GradArena, backward_param_order


def param_table(in_channels: int, n_cls: int, n_filters: int, image_size, transformer_depth: int) -> Dict[str, tuple]:
    """state_dict key -> shape, in the reference's registration order (models/HDenseFormer.py:178-227)."""
    nf, E, g = n_filters, 4 * n_filters, 32
    N = (image_size[0] // 16) * (image_size[1] // 16) * (image_size[2] // 16)
    t: Dict[str, tuple] = {}
    for i in range(in_channels):
        p = f"attns.{i}."
        t[p + "position_embeddings"] = (1, N, E)
        t[p + "patch_embeddings.weight"] = (E, 1, 16, 16, 16)
        t[p + "patch_embeddings.bias"] = (E,)
        for b in range(transformer_depth // 4):
            q = f"{p}blocks.{b}.0."
            for l in range(4):
                r = f"{q}layers.{l}."
                t[r + "0.weight"], t[r + "0.bias"] = (g, E + l * g), (g,)
                t[r + "1.norm.weight"], t[r + "1.norm.bias"] = (g,), (g,)
                t[r + "1.fn.to_qkv.weight"] = (3 * g, g)
                t[r + "1.fn.to_out.0.weight"], t[r + "1.fn.to_out.0.bias"] = (g, g), (g,)
                t[r + "2.norm.weight"], t[r + "2.norm.bias"] = (g,), (g,)
                t[r + "2.fn.net.0.weight"], t[r + "2.fn.net.0.bias"] = (2 * g, g), (2 * g,)
                t[r + "2.fn.net.3.weight"], t[r + "2.fn.net.3.bias"] = (g, 2 * g), (g,)
            t[q + "out_layer.net.0.weight"], t[q + "out_layer.net.0.bias"] = (2 * g, E + 4 * g), (2 * g,)
            t[q + "out_layer.net.3.weight"], t[q + "out_layer.net.3.bias"] = (E, 2 * g), (E,)
    for name, ci, co in (("deep_conv", E * in_channels, 8 * nf), ("up1", 8 * nf, 4 * nf), ("up2", 4 * nf, 2 * nf),
                         ("up3", 2 * nf, nf)):
        t[f"{name}.double_conv.0.weight"], t[f"{name}.double_conv.0.bias"] = (co, ci, 3, 3, 3), (co,)

    def basic(name, ci, co):
        t[f"{name}.conv.weight"] = (co, ci, 3, 3, 3)
        t[f"{name}.norm.weight"], t[f"{name}.norm.bias"] = (co,), (co,)

    def convt(name, ci, co):
        t[f"{name}.weight"], t[f"{name}.bias"] = (ci, co, 3, 3, 3), (co,)

    basic("block_1_1_left", in_channels, nf); basic("block_1_2_left", nf, nf)
    basic("block_2_1_left", nf, 2 * nf); basic("block_2_2_left", 2 * nf, 2 * nf)
    basic("block_3_1_left", 2 * nf, 4 * nf); basic("block_3_2_left", 4 * nf, 4 * nf)
    basic("block_4_1_left", 4 * nf, 8 * nf); basic("block_4_2_left", 8 * nf, 8 * nf)
    convt("upconv_3", 8 * nf, 4 * nf); basic("block_3_1_right", 8 * nf, 4 * nf); basic("block_3_2_right", 4 * nf, 4 * nf)
    convt("upconv_2", 4 * nf, 2 * nf); basic("block_2_1_right", 4 * nf, 2 * nf); basic("block_2_2_right", 2 * nf, 2 * nf)
    convt("upconv_1", 2 * nf, nf); basic("block_1_1_right", 2 * nf, nf); basic("block_1_2_right", nf, nf)
    for name, ci in (("conv1x1", nf), ("conv1x1_d1", 2 * nf), ("conv1x1_d2", 4 * nf), ("conv1x1_d3", 8 * nf)):
        t[f"{name}.weight"], t[f"{name}.bias"] = (n_cls, ci, 1, 1, 1), (n_cls,)
    return t


class _Node(nn.Module):
    """Parameter container; attribute path == reference state_dict key path."""


def _default_init(key: str, shape: tuple) -> torch.Tensor:
    """torch's default initialisers for the corresponding reference layers."""
    if key.endswith("position_embeddings"):
        return torch.zeros(shape)
    if key.endswith("norm.weight"):
        return torch.ones(shape)
    if key.endswith("norm.bias"):
        return torch.zeros(shape)
    return None  # weights / biases handled pairwise below


class _HDFFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, dtype, training, seed, need_grad, *params):
        P = dict(zip(module._keys, params))
        outs, saved = module._engine.forward(P, x, dtype, training, seed, save=need_grad)
        ctx.module, ctx.saved, ctx.P = module, saved, P
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        m = ctx.module
        arena = m._grad_arena()
        # `.grad` tensors are views of the arena.  If some parameter still holds such a view, the arena carries live
        # gradients of an earlier backward that nobody cleared (gradient accumulation, or the model called twice in one
        # autograd graph): keep them aside and add them back, instead of silently wiping them.
        live = any(prm.grad is not None and prm.grad.data_ptr() == arena.views[k].data_ptr() for k, prm in ctx.P.items())
        carried = arena.flat.clone() if live else None
        arena.zero_()
        m._engine.backward(ctx.P, arena.views, ctx.saved, list(gouts), on_grads_ready=None if live else m._on_grads_ready)
        if carried is not None:
            arena.flat.add_(carried)
            m._on_grads_ready(list(arena.views)[-1])      # data-parallel hook: everything is final only now
        ctx.saved = None
        # Hand the gradients over as views of the arena instead of returning them to autograd: AccumulateGrad would clone
        # every view into its own storage (406 device-to-device copies of ~1.5 us each, serial, at the tail of every step:
        # profiles/r1_timeline_v17.txt).  A foreign `.grad` tensor (not an arena view) is accumulated into.
        for k, prm in ctx.P.items():
            if not prm.requires_grad:
                continue
            v = arena.views[k]
            if prm.grad is None:
                prm.grad = v
            elif prm.grad.data_ptr() != v.data_ptr():
                prm.grad.add_(v)
        return (None,) * (6 + len(m._keys))


class HDenseFormer(nn.Module):
    def __init__(self, in_channels, n_cls, n_filters, image_size=(144, 144, 144), transformer_depth=12):
        super().__init__()
        image_size = tuple(image_size) if isinstance(image_size, (tuple, list)) else (image_size,) * 3
        for s in image_size:
            if s % 16 != 0:
                raise ValueError(f"image_size {image_size}: every spatial dim must be a multiple of 16 "
                                 "(patch 16 and 4 x2 up-samplings; the reference fails at ds0+at3 otherwise)")
        self.in_channels, self.n_cls, self.n_filters = in_channels, n_cls, n_filters
        self.image_size, self.transformer_depth = image_size, transformer_depth
        self.compute_dtype = None   # None: follow autocast; or torch.float32 / torch.bfloat16
        self.grad_sync = None       # optional callable(last_ready_key) used by the data-parallel trainer
        table = param_table(in_channels, n_cls, n_filters, image_size, transformer_depth)
        self._keys: List[str] = list(table.keys())
        tensors: Dict[str, torch.Tensor] = {}
        for k, shp in table.items():
            t = _default_init(k, shp)
            if t is None and k.endswith("weight"):
                if "upconv_" in k:
                    fan_in = shp[1] * 27          # torch computes fan_in from dim 1 for ConvTranspose weights
                else:
                    fan_in = int(math.prod(shp[1:]))
                bound = 1.0 / math.sqrt(fan_in)   # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                t = torch.empty(shp).uniform_(-bound, bound)
                bkey = k[:-6] + "bias"
                if bkey in table:
                    tensors[bkey] = torch.empty(table[bkey]).uniform_(-bound, bound)
            if t is not None:
                tensors[k] = t
        for k in self._keys:
            self._register(k, nn.Parameter(tensors[k]))
        self._engine = Engine(Config(in_channels, n_cls, n_filters, image_size, transformer_depth))
        self._arena = None
        self._step = 0
        self._seed_dev = None

    def _register(self, key: str, p: nn.Parameter):
        node = self
        parts = key.split(".")
        for name in parts[:-1]:
            if name not in node._modules:
                node.add_module(name, _Node())
            node = node._modules[name]
        node.register_parameter(parts[-1], p)

    # -- internals used by the autograd function
    def _grad_arena(self) -> GradArena:
        P = dict(self.named_parameters())
        dev = next(iter(P.values())).device
        if self._arena is None or self._arena.flat.device != dev:
            self._arena = GradArena(P, backward_param_order(self._engine.cfg, self._keys))
        return self._arena

    def _on_grads_ready(self, key):
        if self.grad_sync is not None:
            self.grad_sync(key)

    def _resolve_dtype(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            if dt == torch.bfloat16:
                return dt
            raise RuntimeError(f"autocast dtype {dt} is not supported by the B200 path (use bfloat16; fp16+GradScaler "
                               "of the reference is replaced by bf16, see DESIGN.md)")
        return torch.float32

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("hdenseformer_b200 has no CPU path: move the model and the input to a B200 (cuda) device")
        x = x.detach().float().contiguous()
        params = [p for _, p in self.named_parameters()]
        # dropout base seed lives on the device and is advanced by a (graph-capturable) in-place add, so that a
        # captured training step draws fresh masks on every replay
        if self._seed_dev is None or self._seed_dev.device != x.device:
            self._seed_dev = torch.tensor([(torch.initial_seed() * 1000003) & 0x3FFFFFFFFFFFFFFF], dtype=torch.int64,
                                          device=x.device)
        if self.training:
            self._seed_dev.add_(1000003)
        seed = self._seed_dev.clone()     # snapshot: backward of THIS forward must replay the same masks
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        outs = _HDFFunction.apply(self, x, self._resolve_dtype(), self.training, seed, need_grad, *params)
        return list(outs)


def HDenseFormer_32(in_channels, n_cls, image_size, transformer_depth):
    return HDenseFormer(in_channels=in_channels, n_cls=n_cls, image_size=image_size, n_filters=32,
                        transformer_depth=transformer_depth)


def HDenseFormer_16(in_channels, n_cls, image_size, transformer_depth):
    return HDenseFormer(in_channels=in_channels, n_cls=n_cls, image_size=image_size, n_filters=16,
                        transformer_depth=transformer_depth)
