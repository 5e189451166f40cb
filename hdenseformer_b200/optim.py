"""Fused Adam / AdamW for hdenseformer_b200 models (reference: trainer.py:793-840, `_get_optimizer`
with the no-decay grouping of `trainer.py:812-819`: parameters with fewer than 2 dims and `.bias` get weight_decay 0).

One kernel launch updates every parameter from the model's flat gradient arena (csrc/optim.cu); moments are flat fp32
buffers with the arena's offsets.  Learning rate and step count live on the device, so a captured training step can be
replayed while a scheduler changes the rate.  It IS a `torch.optim.Optimizer` (two parameter groups: decayed / not
decayed), so `torch.optim.lr_scheduler.*` and the reference's `PolyLR` (trainer.py:263-264, 1012-1032) accept it: they
write `param_groups[i]['lr']` on the host, `step()` (eager) and `GraphedTrainStep.step()` (before every replay, through
`sync_lr()`) push that value to the device-resident rate.  Both groups share one rate (the reference never uses
per-group rates)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _C
from .ops import _p, _s, ensure_init


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 adamw: bool = False, no_decay_1d: bool = True):
        self.model = model
        self.betas, self.eps, self.adamw = (float(betas[0]), float(betas[1])), float(eps), bool(adamw)
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]
        if not named or not named[0][1].is_cuda:
            raise RuntimeError("FusedAdam needs the model on a CUDA device (hdenseformer_b200 has no CPU path)")
        self.device = named[0][1].device
        ensure_init(named[0][1])
        lib = _C.load()
        arena = model._grad_arena()
        self.arena = arena

        def wd_of(k, p):
            return 0.0 if (no_decay_1d and (p.dim() < 2 or k.endswith(".bias"))) else float(weight_decay)

        decay = [p for k, p in named if wd_of(k, p) > 0]
        no_decay = [p for k, p in named if wd_of(k, p) == 0]
        groups = [{"params": decay, "weight_decay": float(weight_decay)}, {"params": no_decay, "weight_decay": 0.0}]
        groups = [g for g in groups if g["params"]]
        torch.optim.Optimizer.__init__(self, groups, dict(lr=float(lr), betas=self.betas, eps=self.eps,
                                                          weight_decay=float(weight_decay)))
        chunk = lib.hdf_adam_chunk()
        host_tab = ctypes.create_string_buffer(lib.hdf_adam_table_bytes(len(named)))
        chunks = []
        self._keep = []
        for i, (k, p) in enumerate(named):
            if not p.is_contiguous() or p.dtype != torch.float32:
                raise RuntimeError(f"FusedAdam: parameter {k} must be contiguous fp32")
            _C.check(lib.hdf_adam_table_set(host_tab, i, p.data_ptr(), arena.offsets[k], p.numel(), wd_of(k, p)), "adam_table_set")
            chunks += [(i, c) for c in range((p.numel() + chunk - 1) // chunk)]
            self._keep.append(p)
        self.table = torch.frombuffer(bytearray(host_tab.raw), dtype=torch.uint8).to(self.device)
        self.chunks = torch.tensor(chunks, dtype=torch.int32).to(self.device)
        self.nchunks = len(chunks)
        self.m = torch.zeros_like(arena.flat)
        self.v = torch.zeros_like(arena.flat)
        self.hyper = torch.tensor([float(lr), 0.0], dtype=torch.float32, device=self.device)   # {lr, step}
        self._lr_uploaded = float(lr)
        self.grad_scale = 1.0

    # ---- torch.optim-like surface
    def set_lr(self, lr: float):
        """Update the device-resident learning rate (call between graph replays; `param_groups[..]['lr']` is synced too)."""
        for g in self.param_groups:
            g["lr"] = float(lr)
        self.hyper[0:1].fill_(float(lr))
        self._lr_uploaded = float(lr)

    def sync_lr(self):
        """Push a learning rate written to `param_groups[..]['lr']` (LR schedulers do that) to the device copy."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_uploaded:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: learning rate changed while capturing a CUDA graph; call sync_lr() before")
            self.set_lr(lr)

    @torch.no_grad()
    def step(self, closure=None):
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()             # a scheduler wrote param_groups[...]['lr']
        arena = self.model._grad_arena()
        if arena is not self.arena:
            raise RuntimeError("FusedAdam: the model's gradient arena was re-created (device change?); build a new optimizer")
        _C.check(_C.load().hdf_adam_step(_p(self.table), _p(self.chunks), self.nchunks, _p(arena.flat), _p(self.m), _p(self.v),
                                         _p(self.hyper), self.betas[0], self.betas[1], self.eps, int(self.adamw),
                                         float(self.grad_scale), _s()), "adam_step")

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {"m": self.m.clone(), "v": self.v.clone(), "hyper": self.hyper.clone(), "betas": self.betas, "eps": self.eps,
                "adamw": self.adamw}

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.hyper.copy_(sd["hyper"])
        self._lr_uploaded = float(self.hyper[0].item())
        for g in self.param_groups:
            g["lr"] = self._lr_uploaded
