"""Forward / backward executor of the H-DenseFormer 3D graph on libhdf_b200 kernels.

The graph is the reference's (models/HDenseFormer.py:229-255); execution is not: activations are
channels-last (NDHWC) in the compute dtype, every torch.cat is a channel slice of a pre-allocated
buffer that producers write in place, the backward pass is hand-scheduled (no autograd tape per op),
parameter gradients land in one flat fp32 arena (bucketed all-reduce friendly), dropout masks are
recomputed from a counter-based RNG instead of being stored.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, List, Optional

import torch

from . import ops

GROWTH = 32      # models/HDenseFormer.py:79  growth_rate
HEADS = 8        # models/HDenseFormer.py:79
DROP_P = 0.5     # models/HDenseFormer.py:79,105


class GradArena:
    """Flat fp32 gradient buffer with one view per parameter, laid out in the order gradients become
    final during backward (decoder first, transformer last) so that contiguous buckets can be
    all-reduced while the rest of backward is still running."""

    def __init__(self, params: Dict[str, torch.Tensor], order: List[str]):
        self.order = order
        self.offsets = {}
        off = 0
        for k in order:
            self.offsets[k] = off
            off += (params[k].numel() + 31) // 32 * 32   # 128-byte aligned slices
        self.total = off
        dev = params[order[0]].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views = {k: self.flat[self.offsets[k]: self.offsets[k] + params[k].numel()].view(params[k].shape) for k in order}

    def zero_(self):
        self.flat.zero_()


class Config:
    def __init__(self, in_channels, n_cls, n_filters, image_size, transformer_depth):
        self.M = in_channels
        self.n_cls = n_cls
        self.nf = n_filters
        self.image_size = tuple(image_size)
        self.nblocks = transformer_depth // 4
        self.E = 4 * n_filters
        # flat = the 2-D model (reference models/HDenseFormer_2D.py) run as [N, 1, H, W, C] volumes: the depth axis is never
        # pooled / up-sampled / strided, 16 x 16 patches
        self.flat = self.image_size[0] == 1
        self.tok_grid = tuple(1 if (self.flat and i == 0) else s // 16 for i, s in enumerate(image_size))
        self.ntok = self.tok_grid[0] * self.tok_grid[1] * self.tok_grid[2]


def backward_param_order(cfg: Config, keys: List[str]) -> List[str]:
    """Order in which parameter gradients are completed by Engine.backward."""
    def grp(prefix):
        return [k for k in keys if k.startswith(prefix)]
    order: List[str] = []
    for name in ["conv1x1.", "block_1_2_right.", "block_1_1_right.", "upconv_1.", "conv1x1_d1.", "block_2_2_right.",
                 "block_2_1_right.", "upconv_2.", "conv1x1_d2.", "block_3_2_right.", "block_3_1_right.", "upconv_3.",
                 "conv1x1_d3.", "block_4_2_left.", "block_4_1_left.", "block_3_2_left.", "block_3_1_left.",
                 "block_2_2_left.", "block_2_1_left.", "block_1_2_left.", "block_1_1_left.", "up3.", "up2.", "up1.",
                 "deep_conv."]:
        order += grp(name)
    for i in reversed(range(cfg.M)):
        order += grp(f"attns.{i}.")
    assert sorted(order) == sorted(keys), "parameter order table does not cover the state_dict"
    return order


class Ctx:
    """Saved tensors of one forward pass."""
    pass


class Engine:
    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.use_tc = True   # tcgen05 convolution path when available (bf16 only)
        # The transformer branch is ~1100 tiny latency-bound launches; it runs on a second stream next to the big
        # full-resolution convolutions (forward: encoder level 1; backward: the last two encoder blocks).
        # Weight-gradient kernels are off the dependency chain of backward.  They are queued and launched on the main
        # stream AFTER the point where the transformer branch's backward forks onto the side streams, so that ~9 ms of
        # tensor-core work is available to overlap the branch's long chain of small, latency-bound kernels.
        self.defer_wgrad = os.environ.get("HDF_NO_DEFER_WGRAD") is None
        self._deferred = []
        # stream priorities (lower = more urgent; the graphed trainer captures the main path at -1)
        self.fwd_side_priority = int(os.environ.get("HDF_FWD_SIDE_PRIO", "-2"))
        self.bwd_side_priority = int(os.environ.get("HDF_BWD_SIDE_PRIO", "-3"))
        self.wgrad_early = int(os.environ.get("HDF_WGRAD_EARLY", "8"))
        self.tok_wgrad_stream = os.environ.get("HDF_NO_TOK_WGRAD_STREAM") is None
        self.fuse_apply_head = os.environ.get("HDF_NO_APPLY_HEAD") is None
        self._tok_wgrad_stream, self._tok_wgrad_keep = None, []
        self.prepack = os.environ.get("HDF_NO_PREPACK") is None       # conv weights packed up front on a side stream
        self._packed, self._packed_open = {}, False
        self.patch_first = os.environ.get("HDF_NO_PATCH_FIRST") is None   # encoder starts after the patch-embedding GEMMs
        self.use_stem = os.environ.get("HDF_NO_STEM") is None   # im2col + GEMM first layer (bf16 tensor-core path only)
        self.fused_dct = True     # fused post-attention chain kernels (csrc/dct.cu) instead of ~40 single-op launches
        # fused layer-head kernels (Linear_l + LN1 + to_qkv): measured slower than the three single-op launches on B200
        # (34.2 vs 33.1 ms/step), so they stay opt-in
        self.fused_dct_head = os.environ.get("HDF_DCT_HEAD") == "1"
        # bf16 path: tensor-core token kernels (mma.sync Linears / attention) instead of the fp32 SIMT ones
        self.fuse_in_stats = os.environ.get("HDF_NO_FUSED_STATS") is None   # IN statistics from the WS conv epilogue
        self.tok_tc = os.environ.get("HDF_NO_TOK_TC") is None
        self._tok_bf16 = False
        self.use_side_stream = os.environ.get("HDF_NO_SIDE_STREAM") is None
        self._side = {}
        self._keep_alive = None
        self._defer_open = False

    def _side_stream(self, dev, idx=0, phase="f"):
        """Side streams of the transformer branch.  Forward and backward use different stream objects because they want
        different priorities: in the forward pass the token chain is the critical path (the main stream idles at the
        at3 join), in the backward pass it has slack and must not steal SM time from the convolution kernels."""
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), idx, phase)
        if key not in self._side:
            prio = {"f": self.fwd_side_priority, "w": 0}.get(phase, self.bwd_side_priority)
            self._side[key] = torch.cuda.Stream(device=dev, priority=prio)
        return self._side[key]

    # ------------------------------------------------------------------ conv helpers
    def _pack(self, w, Cin, Cout, sci, sco, flip, cin_valid=None):
        """bf16 [27][Cout][Cin] operand tiles of a conv weight; taken from the per-step cache when `_prepack` filled it
        (the consumer's stream then waits for THAT weight's event, not for the whole pre-pack pass)."""
        key = (w.data_ptr(), Cin, Cout, sci, sco, bool(flip), cin_valid)
        return self._pack_cached(key, w, lambda: ops.tc_pack(w, Cin, Cout, sci, sco, flip, cin_valid=cin_valid))

    def _pack_convt(self, w):
        """shared-memory image of a ConvTranspose3d(64 -> 32) weight for the shift-major kernel (csrc/tc_convt.cu)"""
        return self._pack_cached((w.data_ptr(), "convt"), w, lambda: ops.tc_convt_pack(w))

    def _pack_cached(self, key, w, make):
        hit = self._packed.get(key)
        if hit is not None:
            t, ev = hit
            if ev is not None:
                torch.cuda.current_stream(w.device).wait_event(ev)
            return t
        t = make()
        if self._packed_open:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(w.device))
            self._packed[key] = (t, ev)
        return t

    _FWD_USE_ORDER = ("block_1_1_left", "block_1_2_left", "deep_conv", "up1", "up2", "up3", "block_2_1_left", "block_2_2_left",
                      "block_3_1_left", "block_3_2_left", "block_4_1_left", "block_4_2_left", "upconv_3", "block_3_1_right",
                      "block_3_2_right", "upconv_2", "block_2_1_right", "block_2_2_right", "upconv_1", "block_1_1_right",
                      "block_1_2_right")

    def _prepack(self, P, dev, main_stream, need_dgrad):
        """Pack every tensor-core conv weight of the step on a side stream at the start of the forward pass: ~40 small
        kernels leave the main stream's dependency chain (each sat right in front of its convolution).  Weights are packed
        in the order the step uses them (forward layouts in forward order, then the input-gradient layouts in backward
        order) and every one gets its own event, so the first convolutions do not wait for the 2 x 22 MB of the deep layers
        (profiles/r2_timeline_v3.txt: the single end-of-pass event gated the first encoder conv until 1.1 ms)."""
        st = self._side_stream(dev, 99, "p")
        st.wait_stream(main_stream)      # also orders the re-use of last step's buffers behind last step's kernels
        self._packed = {}
        self._packed_open = True
        convs = [(k, w) for k, w in P.items() if w.dim() == 5 and w.shape[2:] == (3, 3, 3)]
        rank = {n: i for i, n in enumerate(self._FWD_USE_ORDER)}
        convs.sort(key=lambda kw: rank.get(kw[0].split(".")[0], len(rank)))
        with torch.cuda.stream(st):
            for k, w in convs:                        # forward layouts
                if k.startswith("upconv_"):          # ConvTranspose3d weight [Cin, Cout, 27]
                    Cin, Cout = w.shape[0], w.shape[1]
                    if ops.tc_convt_supported(Cin, Cout):
                        self._pack_convt(w)
                    elif ops.tc_supported(1, Cin, Cout):
                        self._pack(w, Cin, Cout, Cout * 27, 27, False)
                else:                                 # Conv3d weight [Cout, Cin, 27]
                    Cout, Cin = w.shape[0], w.shape[1]
                    if ops.tc_supported(0, Cin, Cout):
                        self._pack(w, Cin, Cout, 27, Cin * 27, False, cin_valid=Cin)
            if need_dgrad:
                for k, w in reversed(convs):          # input-gradient layouts, in the order backward needs them
                    if k.startswith("upconv_"):
                        Cin, Cout = w.shape[0], w.shape[1]
                        if ops.tc_supported(2, Cout, Cin):
                            self._pack(w, Cout, Cin, 27, Cout * 27, False)
                    else:
                        Cout, Cin = w.shape[0], w.shape[1]
                        if ops.tc_supported(0, Cout, Cin):
                            self._pack(w, Cout, Cin, Cin * 27, 27, True)
        self._packed_open = False
        return None

    def _conv_fwd(self, x, w, bias, out, mode=0):
        """x: [N,D,H,W,Cin] view; w: torch-layout weight; out: [N,Do,Ho,Wo,Cout] view."""
        if mode == 0:    # Conv3d weight [Cout, Cin, 27]
            Cout, Cin = w.shape[0], w.shape[1]
            Cx = x.shape[-1]     # > Cin for the first layer: im2col'ed (stem) or zero-padded to 16 channels
            if Cx != Cin and Cx == ops.stem_kp(Cin):
                return ops.stem_conv_fwd(x, w, out)
            if self.use_tc and x.dtype == torch.bfloat16 and ops.tc_supported(0, Cx, Cout):
                return ops.tc_conv3d_fwd(x, self._pack(w, Cx, Cout, 27, Cin * 27, False, cin_valid=Cin), bias, out)
            assert Cx == Cin
            wp = ops.conv_pack(w, Cin, Cout, 27, Cin * 27, False)
        else:            # ConvTranspose3d weight [Cin, Cout, 27]
            Cin, Cout = w.shape[0], w.shape[1]
            if out.shape[1] == x.shape[1]:
                # flat volume (2-D model): ConvTranspose2d = plane 0 of the 3-D transposed conv of the one-plane input when
                # only the kd = 1 taps are non-zero (out[2j + 0] takes tap kd = 1; the odd plane sees kd = 0 / 2 only)
                tmp = torch.empty((x.shape[0], 2 * x.shape[1], *out.shape[2:4], Cout), dtype=out.dtype, device=out.device)
                self._conv_fwd(x, w, bias, tmp, mode=1)
                out.copy_(tmp[:, ::2])
                return out
            if self.use_tc and x.dtype == torch.bfloat16 and x.shape[-1] == Cin and ops.tc_convt_supported(Cin, Cout):
                return ops.tc_convt_fwd(x, self._pack_convt(w), bias, out)
            if self.use_tc and x.dtype == torch.bfloat16 and ops.tc_supported(1, Cin, Cout):
                return ops.tc_conv3d_fwd(x, self._pack(w, Cin, Cout, Cout * 27, 27, False), bias, out, mode=1)
            wp = ops.conv_pack(w, Cin, Cout, Cout * 27, 27, False)
        return ops.conv3d_fwd(x, wp, bias, out, mode)

    @staticmethod
    def _lift_depth(t):
        """[N, D, H, W, C] view of a flat volume -> dense [N, 2D, H, W, C] with the odd planes zero"""
        z = torch.zeros((t.shape[0], 2 * t.shape[1], *t.shape[2:]), dtype=t.dtype, device=t.device)
        z[:, ::2].copy_(t)
        return z

    def _conv_dgrad(self, dy, w, out, mode=0):
        """input gradient of _conv_fwd(mode): mode 0 -> conv with flipped taps, swapped channels;
        mode 1 (transposed conv) -> strided conv (mode 2)."""
        if mode == 0:
            Cout, Cin = w.shape[0], w.shape[1]
            if self.use_tc and dy.dtype == torch.bfloat16 and ops.tc_supported(0, Cout, Cin):
                # GEMM K = Cout (channels of dy), GEMM N = Cin: packed[tap][ci][co] = w[co][ci][26-tap]
                return ops.tc_conv3d_fwd(dy, self._pack(w, Cout, Cin, Cin * 27, 27, True), None, out)
            wp = ops.conv_pack(w, Cout, Cin, Cin * 27, 27, True)      # packed[tap][co][ci] = w[co][ci][26-tap]
            return ops.conv3d_fwd(dy, wp, None, out, 0)
        Cin, Cout = w.shape[0], w.shape[1]
        if dy.shape[1] == out.shape[1]:
            dy = self._lift_depth(dy)          # flat volume: the odd output plane carries no gradient
        if self.use_tc and dy.dtype == torch.bfloat16 and ops.tc_supported(2, Cout, Cin):
            # strided conv over dy: GEMM K = Cout, N = Cin; packed[tap][n=ci][k=co] = w[ci][co][tap]
            return ops.tc_conv3d_fwd(dy, self._pack(w, Cout, Cin, 27, Cout * 27, False), None, out, mode=2)
        wp = ops.conv_pack(w, Cout, Cin, 27, Cout * 27, False)         # packed[tap][co][ci] = w[ci][co][tap]
        return ops.conv3d_fwd(dy, wp, None, out, 2)

    def _conv_wgrad(self, x, dy, dw, mode=0):
        if mode == 0:    # dw [Cout, Cin, 27]
            Cout, Cin = dw.shape[0], dw.shape[1]
            Cx = x.shape[-1]
            if Cx != Cin and Cx == ops.stem_kp(Cin):
                return ops.stem_conv_wgrad(x, dy, dw)
            if self.use_tc and x.dtype == torch.bfloat16 and ops.tc_wgrad_supported(0, Cx, Cout):
                if Cx == Cin:
                    return ops.tc_conv3d_wgrad(x, dy, dw, 27, Cin * 27, 0)
                tmp = torch.empty((Cout, Cx * 27), dtype=torch.float32, device=dw.device)   # zero-padded input channels
                ops.tc_conv3d_wgrad(x, dy, tmp, 27, Cx * 27, 0)
                return ops.add_rows_f32(dw.view(Cout, Cin * 27), tmp[:, :Cin * 27], False)
            assert Cx == Cin
            ops.conv3d_wgrad(x, dy, dw, 27, Cin * 27, 0)
        else:            # dw [Cin, Cout, 27]
            Cin, Cout = dw.shape[0], dw.shape[1]
            if dy.shape[1] == x.shape[1]:
                dy = self._lift_depth(dy)
            if self.use_tc and x.dtype == torch.bfloat16 and ops.tc_wgrad_supported(1, Cin, Cout):
                return ops.tc_conv3d_wgrad(x, dy, dw, Cout * 27, 27, 1)
            ops.conv3d_wgrad(x, dy, dw, Cout * 27, 27, 1)

    # ------------------------------------------------------------------ BasicConv3d / UpConv
    def _cnr_fwd(self, c: Ctx, name, x, P, out=None, residual=None, affine=True, bias=False, before_apply=None, head=None):
        """conv k3 -> InstanceNorm(+affine) -> ReLU (+ residual).  Saves raw conv output + stats.  head = (weight, bias) of a
        1x1x1 head reading the output: returns (out, logits), fused into the apply pass where the kernel takes the shape."""
        wkey = f"{name}.conv.weight" if affine else f"{name}.double_conv.0.weight"
        w = P[wkey]
        Cout = w.shape[0]
        y = torch.empty((*x.shape[:-1], Cout), dtype=x.dtype, device=x.device)
        cb = P[f"{name}.double_conv.0.bias"] if bias else None
        Cin = w.shape[1]
        if (self.fuse_in_stats and self.use_tc and x.dtype == torch.bfloat16 and x.shape[-1] == Cin and x.shape[0] <= 8
                and ops.tc_ws_supported(0, Cin, Cout)):
            # weight-stationary conv: the InstanceNorm statistics come out of the epilogue (no pass over y)
            mean, rstd = ops.tc_ws_conv3d_fwd_stats(x, self._pack(w, Cin, Cout, 27, Cin * 27, False, cin_valid=Cin), cb, y)
        else:
            self._conv_fwd(x, w, cb, y)
            mean, rstd = ops.instnorm_stats(y)
        if out is None:
            out = torch.empty_like(y)
        g = P[f"{name}.norm.weight"] if affine else None
        b = P[f"{name}.norm.bias"] if affine else None
        if before_apply is not None:
            residual = before_apply()
        setattr(c, name, (x, y, mean, rstd))
        if head is not None:
            if residual is None and self.fuse_apply_head and ops.instnorm_apply_head_supported(y, head[0].shape[0]):
                return out, ops.instnorm_apply_head(y, mean, rstd, g, b, out, head[0], head[1], relu=True)
            ops.instnorm_apply(y, mean, rstd, g, b, out, residual=residual, relu=True)
            return out, ops.head_fwd(out, head[0], head[1])
        ops.instnorm_apply(y, mean, rstd, g, b, out, residual=residual, relu=True)
        return out

    def _cnr_bwd(self, c: Ctx, name, dout, P, G, affine=True, bias=False, need_dx=True):
        x, y, mean, rstd = getattr(c, name)
        wkey = f"{name}.conv.weight" if affine else f"{name}.double_conv.0.weight"
        w = P[wkey]
        g = P[f"{name}.norm.weight"] if affine else None
        b = P[f"{name}.norm.bias"] if affine else None
        dy = ops.instnorm_bwd(dout, y, mean, rstd, g, b, G[f"{name}.norm.weight"] if affine else None,
                              G[f"{name}.norm.bias"] if affine else None, relu=True)
        if self.defer_wgrad and self._defer_open and self._early_left > 0:
            # the first k (HDF_WGRAD_EARLY, default 8) weight gradients of the backward pass start right away on a
            # low-priority stream, next to the bandwidth-bound InstanceNorm passes of the main stream; the rest stay
            # deferred as cover for the transformer backward (k = 0 / 8 / 16: 20.24 / 20.10 / 20.27 ms)
            self._early_left -= 1
            dev = dy.device
            ws = self._side_stream(dev, 9, "w")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            ws.wait_event(ev)
            with torch.cuda.stream(ws):
                self._conv_wgrad(x, dy, G[wkey])
            self._early_keep.append((x, dy))
        elif self.defer_wgrad and self._defer_open:
            self._deferred.append((x, dy, G[wkey], 0, wkey))
        else:
            self._conv_wgrad(x, dy, G[wkey])
        if bias:
            ops.colsum(dy, G[f"{name}.double_conv.0.bias"])
        dx = None
        if need_dx:
            dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
            self._conv_dgrad(dy, w, dx)
        return dx

    # ------------------------------------------------------------------ DCT block
    def _dct_block_fwd(self, P, pre, F, Cin_tok, R, B, training, seed, ids, saved):
        """F: [R, E+128] feature buffer whose first E columns hold the block input.  Returns o1 (the
        hidden of out_layer); the caller runs the last Linear so it can write into the next buffer."""
        E = self.cfg.E
        p = DROP_P if training else 0.0
        dev = F.device
        f32 = torch.float32
        layers = []
        for l in range(4):
            q = f"{pre}layers.{l}."
            Cl = E + GROWTH * l
            if self.tok_tc and self._tok_bf16:
                # bf16 path: the whole inner layer in two tensor-core kernels (csrc/tok_tc.cu)
                h0, n1, m1, r1, qkv = ops.tok_a_fwd(F, Cl, P, q)
                ida, idb, idc, idd, ide = ids(), ids(), ids(), ids(), ids()
                o, lse, sv = ops.tok_c_fwd(qkv, h0, P, q, F[:, Cl:Cl + GROWTH], B, R // B, (GROWTH // HEADS) ** -0.5, p, seed,
                                           (ida, idb, idc, idd, ide))
                layers.append(dict(h0=h0, n1=n1, m1=m1, r1=r1, qkv=qkv, o=o, lse=lse, sv=sv, ids=(ida, idb, idc, idd, ide)))
                continue
            if self.fused_dct_head:
                h0, n1, m1, r1, qkv = ops.dct_a_fwd(F, Cl, P, q)
            else:
                h0 = torch.empty((R, GROWTH), dtype=f32, device=dev)
                ops.gemm(F[:, :Cl], P[q + "0.weight"], True, h0, bias=P[q + "0.bias"])
                n1, m1, r1 = ops.layernorm_fwd(h0, P[q + "1.norm.weight"], P[q + "1.norm.bias"])
                qkv = torch.empty((R, 3 * GROWTH), dtype=f32, device=dev)
                ops.gemm(n1, P[q + "1.fn.to_qkv.weight"], True, qkv)
            o, lse = ops.attention_fwd(qkv, B, R // B, HEADS, (GROWTH // HEADS) ** -0.5)
            ida, idb, idc, idd, ide = ids(), ids(), ids(), ids(), ids()
            if self.fused_dct:
                sv = ops.dct_c_fwd(o, h0, P, q, F[:, Cl:Cl + GROWTH], p, seed, (ida, idb, idc, idd, ide))
                layers.append(dict(h0=h0, n1=n1, m1=m1, r1=r1, qkv=qkv, o=o, lse=lse, sv=sv, ids=(ida, idb, idc, idd, ide)))
                continue
            h1 = torch.empty((R, GROWTH), dtype=f32, device=dev)
            ops.gemm(o, P[q + "1.fn.to_out.0.weight"], True, h1, bias=P[q + "1.fn.to_out.0.bias"], residual=h0, p=p, seed=seed,
                     call_id=ida)
            n2, m2, r2 = ops.layernorm_fwd(h1, P[q + "2.norm.weight"], P[q + "2.norm.bias"])
            z1 = torch.empty((R, 2 * GROWTH), dtype=f32, device=dev)
            f1 = torch.empty_like(z1)
            ops.gemm(n2, P[q + "2.fn.net.0.weight"], True, f1, bias=P[q + "2.fn.net.0.bias"], pre=z1, act=1, p=p, seed=seed,
                     call_id=idb)
            h2 = torch.empty((R, GROWTH), dtype=f32, device=dev)
            ops.gemm(f1, P[q + "2.fn.net.3.weight"], True, h2, bias=P[q + "2.fn.net.3.bias"], residual=h1, p=p, seed=seed,
                     call_id=idc)
            n3, m3, r3 = ops.layernorm_fwd(h2, P[q + "2.norm.weight"], P[q + "2.norm.bias"])
            z1b = torch.empty((R, 2 * GROWTH), dtype=f32, device=dev)
            g1 = torch.empty_like(z1b)
            ops.gemm(n3, P[q + "2.fn.net.0.weight"], True, g1, bias=P[q + "2.fn.net.0.bias"], pre=z1b, act=1, p=p, seed=seed,
                     call_id=idd)
            ops.gemm(g1, P[q + "2.fn.net.3.weight"], True, F[:, Cl:Cl + GROWTH], bias=P[q + "2.fn.net.3.bias"], p=p, seed=seed,
                     call_id=ide)
            layers.append(dict(h0=h0, n1=n1, m1=m1, r1=r1, qkv=qkv, o=o, lse=lse, h1=h1, n2=n2, m2=m2, r2=r2, z1=z1, f1=f1,
                               h2=h2, n3=n3, m3=m3, r3=r3, z1b=z1b, g1=g1, ids=(ida, idb, idc, idd, ide)))
        zo = torch.empty((R, 2 * GROWTH), dtype=f32, device=dev)
        o1 = torch.empty_like(zo)
        idf = ids()
        ops.gemm(F, P[pre + "out_layer.net.0.weight"], True, o1, bias=P[pre + "out_layer.net.0.bias"], pre=zo, act=1, p=p,
                 seed=seed, call_id=idf)
        saved.update(F=F, layers=layers, zo=zo, o1=o1, idf=idf)
        return o1

    def _dct_c_bwd_unfused(self, P, G, q, s, dg2, R, p, seed):
        """reference composition of the chain backward from single-op kernels (kept for A/B testing)"""
        dev = dg2.device
        f32 = torch.float32
        ida, idb, idc, idd, ide = s["ids"]
        W1, W2 = P[q + "2.fn.net.0.weight"], P[q + "2.fn.net.3.weight"]
        # features.append(ff(LN(h2)))
        dzz = ops.act_dropout_bwd(dg2, None, 0, p, seed, ide)
        ops.gemm_at_b(dzz, s["g1"], G[q + "2.fn.net.3.weight"])
        ops.colsum(dzz, G[q + "2.fn.net.3.bias"], accumulate=True)
        dg1 = torch.empty((R, 2 * GROWTH), dtype=f32, device=dev)
        ops.gemm(dzz, W2, False, dg1)
        dz1b = ops.act_dropout_bwd(dg1, s["z1b"], 1, p, seed, idd)
        ops.gemm_at_b(dz1b, s["n3"], G[q + "2.fn.net.0.weight"])
        ops.colsum(dz1b, G[q + "2.fn.net.0.bias"], accumulate=True)
        dn3 = torch.empty((R, GROWTH), dtype=f32, device=dev)
        ops.gemm(dz1b, W1, False, dn3)
        dh = torch.empty((R, GROWTH), dtype=f32, device=dev)      # running grad of the residual stream
        ops.layernorm_bwd(dn3, s["h2"], s["m3"], s["r3"], P[q + "2.norm.weight"], dh, False, G[q + "2.norm.weight"],
                          G[q + "2.norm.bias"])
        # h2 = drop(f1 W2^T + b2) + h1
        dzz2 = ops.act_dropout_bwd(dh, None, 0, p, seed, idc)
        ops.gemm_at_b(dzz2, s["f1"], G[q + "2.fn.net.3.weight"])
        ops.colsum(dzz2, G[q + "2.fn.net.3.bias"], accumulate=True)
        df1 = torch.empty((R, 2 * GROWTH), dtype=f32, device=dev)
        ops.gemm(dzz2, W2, False, df1)
        dz1 = ops.act_dropout_bwd(df1, s["z1"], 1, p, seed, idb)
        ops.gemm_at_b(dz1, s["n2"], G[q + "2.fn.net.0.weight"])
        ops.colsum(dz1, G[q + "2.fn.net.0.bias"], accumulate=True)
        dn2 = torch.empty((R, GROWTH), dtype=f32, device=dev)
        ops.gemm(dz1, W1, False, dn2)
        ops.layernorm_bwd(dn2, s["h1"], s["m2"], s["r2"], P[q + "2.norm.weight"], dh, True, G[q + "2.norm.weight"],
                          G[q + "2.norm.bias"])
        # h1 = drop(o Wo^T + bo) + h0
        dzo_ = ops.act_dropout_bwd(dh, None, 0, p, seed, ida)
        ops.gemm_at_b(dzo_, s["o"], G[q + "1.fn.to_out.0.weight"])
        ops.colsum(dzo_, G[q + "1.fn.to_out.0.bias"], accumulate=True)
        do = torch.empty((R, GROWTH), dtype=f32, device=dev)
        ops.gemm(dzo_, P[q + "1.fn.to_out.0.weight"], False, do)
        return do, dh

    def _tok_off_chain(self, fn, *keep):
        """Run a weight-gradient launch of the token branch off its dependency chain: on a second stream that waits for
        the chain's current position.  The chain is ~100 small dependent kernels per modality and ends the step; the
        Linear weight / bias gradients hang off it as leaves (HDF_NO_TOK_WGRAD_STREAM=1 keeps them in line)."""
        ws = self._tok_wgrad_stream
        if ws is None:
            return fn()
        cur = torch.cuda.current_stream(keep[0].device)
        ev = torch.cuda.Event()
        ev.record(cur)
        ws.wait_event(ev)
        with torch.cuda.stream(ws):
            fn()
        self._tok_wgrad_keep.append(keep)       # operands stay alive until the join at the end of the branch

    def _dct_block_bwd(self, P, G, pre, saved, d_o1, R, B, training, seed):
        """d_o1: grad wrt out_layer hidden (after GELU+dropout).  Returns dX = grad wrt block input [R,E] view."""
        E = self.cfg.E
        p = DROP_P if training else 0.0
        F = saved["F"]
        dev = F.device
        f32 = torch.float32
        dzo = ops.act_dropout_bwd(d_o1, saved["zo"], 1, p, seed, saved["idf"])
        self._tok_off_chain(lambda: (ops.gemm_at_b(dzo, F, G[pre + "out_layer.net.0.weight"]),
                                     ops.colsum(dzo, G[pre + "out_layer.net.0.bias"], accumulate=True)), dzo, F)
        dF = torch.empty_like(F)
        ops.gemm(dzo, P[pre + "out_layer.net.0.weight"], False, dF)
        scale = (GROWTH // HEADS) ** -0.5
        for l in reversed(range(4)):
            q = f"{pre}layers.{l}."
            s = saved["layers"][l]
            ida, idb, idc, idd, ide = s["ids"]
            Cl = E + GROWTH * l
            if self.fused_dct:
                do, dh = ops.dct_c_bwd(dF[:, Cl:Cl + GROWTH], s["o"], s["sv"], P, G, q, p, seed, s["ids"])
            else:
                do, dh = self._dct_c_bwd_unfused(P, G, q, s, dF[:, Cl:Cl + GROWTH], R, p, seed)
            dqkv = ops.attention_bwd(s["qkv"], s["o"], do, s["lse"], B, R // B, HEADS, scale)
            if self.fused_dct_head:
                ops.dct_a_bwd(dqkv, dh, s, F, Cl, dF, P, G, q)
                continue
            self._tok_off_chain(lambda dqkv=dqkv, s=s, q=q: ops.gemm_at_b(dqkv, s["n1"], G[q + "1.fn.to_qkv.weight"]), dqkv, s["n1"])
            dn1 = torch.empty((R, GROWTH), dtype=f32, device=dev)
            ops.gemm(dqkv, P[q + "1.fn.to_qkv.weight"], False, dn1)
            ops.layernorm_bwd(dn1, s["h0"], s["m1"], s["r1"], P[q + "1.norm.weight"], dh, True, G[q + "1.norm.weight"],
                              G[q + "1.norm.bias"])
            # h0 = F[:, :Cl] W_l^T + b_l
            self._tok_off_chain(lambda dh=dh, q=q, Cl=Cl: (ops.gemm_at_b(dh, F[:, :Cl], G[q + "0.weight"]),
                                                          ops.colsum(dh, G[q + "0.bias"], accumulate=True)), dh, F)
            ops.gemm(dh, P[q + "0.weight"], False, dF[:, :Cl], accumulate=True)
        return dF[:, :E]

    # ------------------------------------------------------------------ forward
    def forward(self, P: Dict[str, torch.Tensor], x: torch.Tensor, dtype: torch.dtype, training: bool, seed: int,
                save: bool = True):
        cfg = self.cfg
        ops.ensure_init(x)
        assert x.dtype == torch.float32 and x.is_contiguous()
        B, M, D, H, W = x.shape
        assert M == cfg.M and (D, H, W) == cfg.image_size, \
            f"input {tuple(x.shape)} does not match in_channels={cfg.M}, image_size={cfg.image_size}"
        nf, E = cfg.nf, cfg.E
        dev = x.device
        c = Ctx()
        c.x, c.dtype, c.training, c.seed, c.B = x, dtype, training, seed, B
        self._tok_bf16 = dtype == torch.bfloat16
        counter = [0]

        def ids():
            counter[0] += 1
            return counter[0]

        def empty(shape, dt=dtype):
            return torch.empty(shape, dtype=dt, device=dev)

        main_stream = torch.cuda.current_stream(dev)
        side = self._side_stream(dev) if self.use_side_stream else None
        pack_ev = None
        self._packed, self._packed_open = {}, False
        if side is not None and self.prepack and self.use_tc and dtype == torch.bfloat16:
            pack_ev = self._prepack(P, dev, main_stream, need_dgrad=save)
        if side is not None:
            side.wait_stream(main_stream)                                   # fork
        branch_ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
        branch_ctx.__enter__()
        fork_ev = None
        if side is not None:
            fork_ev = torch.cuda.Event()
            fork_ev.record(side)             # modality streams fork from HERE, not from the end of modality 0's work
        # ---------------- transformer branches (fp32 tokens), one per modality
        d16 = cfg.tok_grid
        ntok = cfg.ntok
        R = B * ntok
        FW = E + 4 * GROWTH
        attnall = empty((B, *d16, E * M))
        c.tr = []
        p = DROP_P if training else 0.0
        mod_streams = []
        pe_events = []
        for i in range(M):
            pre = f"attns.{i}."
            tr = dict(blocks=[])
            # modality branches are independent: branch i > 0 gets its own stream (forked from / joined into branch 0)
            ms = self._side_stream(dev, i) if (side is not None and i > 0) else None
            if ms is not None:
                ms.wait_event(fork_ev)
                mod_ctx = torch.cuda.stream(ms)
                mod_ctx.__enter__()
            F = empty((R, FW), torch.float32)
            tr["pe_id"] = ids()
            ops.patch_embed_fwd(x, i, P[pre + "patch_embeddings.weight"], P[pre + "patch_embeddings.bias"],
                                P[pre + "position_embeddings"], F[:, :E], p, seed, tr["pe_id"],
                                tensor_cores=self.tok_tc and self._tok_bf16)
            if side is not None and self.patch_first:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                pe_events.append(ev)
            tok = None
            for b in range(cfg.nblocks):
                bp = f"{pre}blocks.{b}.0."
                saved = {}
                o1 = self._dct_block_fwd(P, bp, F, E, R, B, training, seed, ids, saved)
                saved["idg"] = ids()
                if b + 1 < cfg.nblocks:
                    Fn = empty((R, FW), torch.float32)
                    dst = Fn[:, :E]
                else:
                    Fn = None
                    tok = empty((R, E), torch.float32)
                    dst = tok
                ops.gemm(o1, P[bp + "out_layer.net.3.weight"], True, dst, bias=P[bp + "out_layer.net.3.bias"], p=p, seed=seed,
                         call_id=saved["idg"])
                tr["blocks"].append(saved)
                F = Fn
            if cfg.nblocks == 0:
                tok = F[:, :E]
            ops.cast_from_f32(tok, attnall.view(R, E * M)[:, i * E:(i + 1) * E])
            c.tr.append(tr)
            if ms is not None:
                mod_ctx.__exit__(None, None, None)
                mod_streams.append(ms)
        for ms in mod_streams:
            side.wait_stream(ms)
        c.attnall = attnall

        # ---------------- up-sampling path of the transformer features (UpConv x4)
        def upconv(name, xin):
            a = self._cnr_fwd(c, name, xin, P, affine=False, bias=True)
            up = empty((B, (1 if cfg.flat else 2) * a.shape[1], 2 * a.shape[2], 2 * a.shape[3], a.shape[4]))
            return ops.upsample2_fwd(a, up)

        attnout = upconv("deep_conv", attnall)
        at1 = upconv("up1", attnout)
        at2 = upconv("up2", at1)
        at3 = upconv("up3", at2)
        branch_ctx.__exit__(None, None, None)
        self._keep_alive = (attnall, attnout, at1, at2, at3)   # produced on the side stream, consumed on the main one

        def join_at3():
            if side is not None:
                main_stream.wait_stream(side)                                   # join before the first consumer
            return at3

        # ---------------- encoder; skip tensors are written straight into the decoder concat buffers
        # The token chain is the forward critical path (the main stream idles ~1 ms at join_at3): let the patch-embedding
        # GEMMs have the GPU to themselves instead of queueing behind the 41 k blocks of the first encoder kernels.
        for ev in pe_events:
            main_stream.wait_event(ev)
        if self.use_tc and self.use_stem and dtype == torch.bfloat16 and ops.stem_fused_supported(M, nf):
            xcl = ops.StemInput(x)          # first conv gathers its taps from the fp32 volume itself (csrc/stem_tc.cu)
        elif self.use_tc and self.use_stem and dtype == torch.bfloat16 and ops.stem_supported(M, nf):
            xcl = ops.stem_im2col(x)        # first conv = one GEMM over the gathered taps (csrc/tc_conv.cu, stem path)
        else:
            pad = 16 if (self.use_tc and dtype == torch.bfloat16 and M < 16 and ops.tc_supported(0, 16, nf)) else 0
            xcl = ops.ncdhw_to_cl(x, dtype, pad_to=pad)
        c.xcl = xcl
        cat1 = empty((B, D, H, W, 2 * nf))
        dl = (lambda l: D) if cfg.flat else (lambda l: D >> l)      # depth of pyramid level l
        cat2 = empty((B, dl(1), H // 2, W // 2, 4 * nf))
        cat3 = empty((B, dl(2), H // 4, W // 4, 8 * nf))
        c.cat1, c.cat2, c.cat3 = cat1, cat2, cat3
        a = self._cnr_fwd(c, "block_1_1_left", xcl, P)
        ds0 = self._cnr_fwd(c, "block_1_2_left", a, P, out=cat1[..., nf:], before_apply=join_at3)
        p1 = ops.maxpool2_fwd(ds0, empty((B, dl(1), H // 2, W // 2, nf)))
        a = self._cnr_fwd(c, "block_2_1_left", p1, P)
        ds1 = self._cnr_fwd(c, "block_2_2_left", a, P, out=cat2[..., 2 * nf:], residual=at2)
        p2 = ops.maxpool2_fwd(ds1, empty((B, dl(2), H // 4, W // 4, 2 * nf)))
        a = self._cnr_fwd(c, "block_3_1_left", p2, P)
        ds2 = self._cnr_fwd(c, "block_3_2_left", a, P, out=cat3[..., 4 * nf:], residual=at1)
        p3 = ops.maxpool2_fwd(ds2, empty((B, dl(3), H // 8, W // 8, 4 * nf)))
        a = self._cnr_fwd(c, "block_4_1_left", p3, P)
        x4 = self._cnr_fwd(c, "block_4_2_left", a, P, residual=attnout)
        c.x4 = x4

        # ---------------- decoder with deep-supervision heads
        out3 = ops.head_fwd(x4, P["conv1x1_d3.weight"], P["conv1x1_d3.bias"])
        self._conv_fwd(x4, P["upconv_3.weight"], P["upconv_3.bias"], cat3[..., :4 * nf], mode=1)
        a = self._cnr_fwd(c, "block_3_1_right", cat3, P)
        a32, out2 = self._cnr_fwd(c, "block_3_2_right", a, P, head=(P["conv1x1_d2.weight"], P["conv1x1_d2.bias"]))
        self._conv_fwd(a32, P["upconv_2.weight"], P["upconv_2.bias"], cat2[..., :2 * nf], mode=1)
        a = self._cnr_fwd(c, "block_2_1_right", cat2, P)
        a22, out1 = self._cnr_fwd(c, "block_2_2_right", a, P, head=(P["conv1x1_d1.weight"], P["conv1x1_d1.bias"]))
        self._conv_fwd(a22, P["upconv_1.weight"], P["upconv_1.bias"], cat1[..., :nf], mode=1)
        a = self._cnr_fwd(c, "block_1_1_right", cat1, P)
        a12, out0 = self._cnr_fwd(c, "block_1_2_right", a, P, head=(P["conv1x1.weight"], P["conv1x1.bias"]))
        c.a32, c.a22, c.a12 = a32, a22, a12
        return [out0, out1, out2, out3], (c if save else None)

    # ------------------------------------------------------------------ backward
    def backward(self, P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor], c: Ctx, gouts: List[Optional[torch.Tensor]],
                 on_grads_ready=None):
        """Fills G (views of the GradArena) with parameter gradients.  `on_grads_ready(last_key)` is called
        whenever all gradients up to and including `last_key` (arena order) are final."""
        cfg = self.cfg
        nf, E, B = cfg.nf, cfg.E, c.B
        dt = c.dtype
        dev = c.x.device
        notify = on_grads_ready or (lambda k: None)

        def empty(shape, dtype=dt):
            return torch.empty(shape, dtype=dtype, device=dev)

        def gout(i, like_shape):
            g = gouts[i]
            if g is None:
                return torch.zeros(like_shape, dtype=dt, device=dev)
            return g.to(dt).contiguous()

        def head_bwd(name, g, a, da, accumulate_da):
            ops.head_bwd(g, a, P[name + ".weight"].view(cfg.n_cls, -1), da, G[name + ".weight"], G[name + ".bias"], accumulate_da)

        def convt_bwd(name, du, a_in):
            """du: grad of the transposed-conv output (a channel-slice view); returns fresh grad wrt its input"""
            da = empty(a_in.shape)
            self._conv_dgrad(du, P[name + ".weight"], da, mode=1)
            if self.defer_wgrad and self._defer_open:
                self._deferred.append((a_in, du, G[name + ".weight"], 1, name + ".weight"))
            else:
                self._conv_wgrad(a_in, du, G[name + ".weight"], mode=1)
            ops.colsum(du, G[name + ".bias"])
            return da

        D, H, W = cfg.image_size
        dl = (lambda l: D) if cfg.flat else (lambda l: D >> l)
        self._deferred = []
        self._defer_open = self.use_side_stream and self.defer_wgrad   # only useful with a side branch to overlap
        self._early_left = self.wgrad_early if self._defer_open else 0
        self._early_keep = []
        # ---- level 0 (full resolution)
        dA = empty(c.a12.shape)
        head_bwd("conv1x1", gout(0, (B, cfg.n_cls, D, H, W)), c.a12, dA, False)
        dA = self._cnr_bwd(c, "block_1_2_right", dA, P, G)
        dcat1 = self._cnr_bwd(c, "block_1_1_right", dA, P, G)
        dA = convt_bwd("upconv_1", dcat1[..., :nf], c.a22)
        head_bwd("conv1x1_d1", gout(1, (B, cfg.n_cls, dl(1), H // 2, W // 2)), c.a22, dA, True)
        if not self._defer_open:
            notify("conv1x1_d1.bias")
        # ---- level 1
        dA = self._cnr_bwd(c, "block_2_2_right", dA, P, G)
        dcat2 = self._cnr_bwd(c, "block_2_1_right", dA, P, G)
        dA = convt_bwd("upconv_2", dcat2[..., :2 * nf], c.a32)
        head_bwd("conv1x1_d2", gout(2, (B, cfg.n_cls, dl(2), H // 4, W // 4)), c.a32, dA, True)
        if not self._defer_open:
            notify("conv1x1_d2.bias")
        # ---- level 2
        dA = self._cnr_bwd(c, "block_3_2_right", dA, P, G)
        dcat3 = self._cnr_bwd(c, "block_3_1_right", dA, P, G)
        dx4 = convt_bwd("upconv_3", dcat3[..., :4 * nf], c.x4)
        head_bwd("conv1x1_d3", gout(3, (B, cfg.n_cls, dl(3), H // 8, W // 8)), c.x4, dx4, True)
        if not self._defer_open:
            notify("conv1x1_d3.bias")
        # ---- bottleneck + encoder (dx4 is also the gradient of attnout through the residual add)
        dA = self._cnr_bwd(c, "block_4_2_left", dx4, P, G)
        dp3 = self._cnr_bwd(c, "block_4_1_left", dA, P, G)
        dds2 = dcat3[..., 4 * nf:]
        ops.maxpool2_bwd(c.cat3[..., 4 * nf:], dp3, dds2, True)
        dA = self._cnr_bwd(c, "block_3_2_left", dds2, P, G)
        dp2 = self._cnr_bwd(c, "block_3_1_left", dA, P, G)
        dds1 = dcat2[..., 2 * nf:]
        ops.maxpool2_bwd(c.cat2[..., 2 * nf:], dp2, dds1, True)
        if not self._defer_open:
            notify("block_3_1_left.norm.bias")
        dA = self._cnr_bwd(c, "block_2_2_left", dds1, P, G)
        dp1 = self._cnr_bwd(c, "block_2_1_left", dA, P, G)
        dds0 = dcat1[..., nf:]
        ops.maxpool2_bwd(c.cat1[..., nf:], dp1, dds0, True)
        main_stream = torch.cuda.current_stream(dev)
        side = self._side_stream(dev, 0, "b") if self.use_side_stream else None

        # ---- transformer-feature up path: at3 <- up3 <- at2 <- up2 <- at1 <- up1 <- attnout <- deep_conv <- attnall.
        # It is the head of the longest remaining dependency chain (the transformer backward hangs off it), and its
        # convolutions cannot share SMs with the other persistent conv kernels, so it goes FIRST on the main stream.
        def upconv_bwd(name, dup, extra):
            """dup: grad wrt the upsampled output; extra: additional grad wrt this UpConv's *input* (or None)"""
            xin, y, _, _ = getattr(c, name)
            da = empty(y.shape)
            ops.upsample2_bwd(dup, da, False)
            dx = self._cnr_bwd(c, name, da, P, G, affine=False, bias=True)
            if extra is not None:
                ops.add_(dx, extra)
            return dx

        dat2 = upconv_bwd("up3", dds0, dds1)
        dat1 = upconv_bwd("up2", dat2, dds2)
        dattnout = upconv_bwd("up1", dat1, dx4)
        dattnall = upconv_bwd("deep_conv", dattnout, None)

        fork_ev = None
        if side is not None:
            side.wait_stream(main_stream)       # fork: the transformer backward (small kernels) runs beside the window below
            fork_ev = torch.cuda.Event()
            fork_ev.record(side)
        # ---- overlap window on the main stream: last two encoder blocks + every deferred weight gradient
        self._defer_open = False
        dA = self._cnr_bwd(c, "block_1_2_left", dds0, P, G)
        self._cnr_bwd(c, "block_1_1_left", dA, P, G, need_dx=False)
        # Gradient buckets for the data-parallel all-reduce: everything except the transformer branches is final once the
        # deferred weight gradients have run, and they complete in arena order, so after the j-th of them the arena
        # prefix up to (not including) the next deferred weight is final.  The all-reduces issued here (NCCL stream,
        # ordered behind the main stream) overlap the remaining weight gradients and the transformer backward.
        order = list(G.keys())
        pos = {k: i for i, k in enumerate(order)}
        if self._early_keep:
            main_stream.wait_stream(self._side_stream(dev, 9, "w"))     # early weight gradients are final before any bucket leaves
            self._early_keep = []
        for j, (wx, wdy, wg, wmode, wkey) in enumerate(self._deferred):
            self._conv_wgrad(wx, wdy, wg, mode=wmode)
            if side is not None:
                nxt = self._deferred[j + 1][4] if j + 1 < len(self._deferred) else None
                notify(order[pos[nxt] - 1] if nxt is not None else "deep_conv.double_conv.0.bias")
        if not self._deferred and side is not None:
            notify("deep_conv.double_conv.0.bias")
        self._deferred = []
        if side is None:
            notify("deep_conv.double_conv.0.bias")
        branch_ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
        branch_ctx.__enter__()

        # ---- transformer branches
        R = B * cfg.ntok
        p = DROP_P if c.training else 0.0
        mod_streams = []
        for i in reversed(range(cfg.M)):
            pre = f"attns.{i}."
            tr = c.tr[i]
            ms = self._side_stream(dev, i, "b") if (side is not None and i > 0) else None
            if ms is not None:
                ms.wait_event(fork_ev)
                mod_ctx = torch.cuda.stream(ms)
                mod_ctx.__enter__()
            dtok = empty((R, E), torch.float32)
            ops.cast_to_f32(dattnall.view(R, E * cfg.M)[:, i * E:(i + 1) * E], dtok)
            self._tok_wgrad_stream = self._side_stream(dev, 20 + i, "b") if (side is not None and self.tok_wgrad_stream) else None
            for b in reversed(range(cfg.nblocks)):
                bp = f"{pre}blocks.{b}.0."
                saved = tr["blocks"][b]
                dz = ops.act_dropout_bwd(dtok, None, 0, p, c.seed, saved["idg"])
                self._tok_off_chain(lambda dz=dz, saved=saved, bp=bp: (ops.gemm_at_b(dz, saved["o1"], G[bp + "out_layer.net.3.weight"]),
                                                                       ops.colsum(dz, G[bp + "out_layer.net.3.bias"], accumulate=True)),
                                    dz, saved["o1"])
                d_o1 = empty((R, 2 * GROWTH), torch.float32)
                ops.gemm(dz, P[bp + "out_layer.net.3.weight"], False, d_o1)
                dtok = self._dct_block_bwd(P, G, bp, saved, d_o1, R, B, c.training, c.seed)
            # patch embedding: tok = drop(conv(img) + bias + pos)
            dpe = ops.act_dropout_bwd(dtok, None, 0, p, c.seed, tr["pe_id"])
            # leaves of the chain: run beside the other modality's chain instead of at the tail of this one
            self._tok_off_chain(lambda dpe=dpe, pre=pre, i=i: (ops.posemb_grad(dpe, G[pre + "position_embeddings"], B, cfg.ntok, E),
                                                               ops.colsum(dpe, G[pre + "patch_embeddings.bias"]),
                                                               ops.patch_embed_wgrad(c.x, i, dpe, G[pre + "patch_embeddings.weight"])),
                                dpe)
            if self._tok_wgrad_stream is not None:
                torch.cuda.current_stream(dev).wait_stream(self._tok_wgrad_stream)
                self._tok_wgrad_stream, self._tok_wgrad_keep = None, []
            if side is None:
                notify([k for k in G if k.startswith(pre)][-1])
            if ms is not None:
                mod_ctx.__exit__(None, None, None)
                mod_streams.append(ms)
        for ms in mod_streams:
            side.wait_stream(ms)
        branch_ctx.__exit__(None, None, None)
        if side is not None:
            main_stream.wait_stream(side)       # join; the remaining gradient buckets are released together
            notify([k for k in G][-1])
