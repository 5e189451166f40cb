"""Thin typed wrappers over the C ABI: torch tensors in, torch tensors out.  torch is used for
device memory and streams only; every arithmetic op here is a kernel of libhdf_b200.so."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _C

_DT = {torch.float32: _C.F32, torch.bfloat16: _C.BF16}


def _lib():
    return _C.load()


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    # kernels launch on the CURRENT device's current stream (_s, Workspace): an operand on another GPU must fail loudly,
    # not run on the wrong device (one process per GPU is the supported layout; use torch.cuda.device(...) otherwise)
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise _C.HDFError(f"operand on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}")
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ld(t: torch.Tensor) -> int:
    """channel stride (elements) of a channels-last view [..., C]"""
    assert t.stride(-1) == 1, "innermost (channel) dim must be contiguous"
    ld = t.stride(-2)
    # all leading dims must be densely packed on top of ld
    exp = ld
    for i in range(t.dim() - 2, -1, -1):
        assert t.shape[i] == 1 or t.stride(i) == exp, f"view is not row-dense: shape {tuple(t.shape)} stride {t.stride()}"
        exp *= t.shape[i]
    return ld


class Workspace:
    """One growable scratch buffer per (device, stream).  A buffer that is outgrown is retired, never freed: a captured
    CUDA graph (GraphedTrainStep, graphed sliding window) may have its address baked into kernel arguments and tensor maps."""
    _bufs = {}
    _retired = []

    @classmethod
    def get(cls, nbytes: int) -> torch.Tensor:
        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                cls._retired.append(buf)
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device="cuda")
            cls._bufs[key] = buf
        return buf


def ensure_init(t: torch.Tensor):
    if not t.is_cuda:
        raise _C.HDFError("hdenseformer_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device")
    _C.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


# ----------------------------------------------------------------------------- conv
def conv_pack(w: torch.Tensor, A: int, B: int, stride_a: int, stride_b: int, flip: bool) -> torch.Tensor:
    out = torch.empty((27, A, B), dtype=torch.float32, device=w.device)
    _C.check(_lib().hdf_conv_pack_weights(_p(w), _p(out), A, B, stride_a, stride_b, int(flip), _s()), "conv_pack")
    return out


def conv3d_fwd(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, mode: int = 0):
    N, Do, Ho, Wo, Cout = out.shape
    Cin = x.shape[-1]
    _C.check(_lib().hdf_conv3d_fwd(_DT[x.dtype], mode, _p(x), _ld(x), _p(wp), _p(bias), _p(out), _ld(out), N, Do, Ho, Wo, Cin,
                                   Cout, _s()), "conv3d_fwd")
    return out


def conv3d_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, stride_ci: int, stride_co: int, mode: int = 0,
                 accumulate: bool = False):
    N, Do, Ho, Wo, Cout = dy.shape
    Cin = x.shape[-1]
    nb = _lib().hdf_conv3d_wgrad_workspace(N, Do, Ho, Wo, Cin, Cout)
    ws = Workspace.get(nb)
    _C.check(_lib().hdf_conv3d_wgrad(_DT[x.dtype], mode, _p(x), _ld(x), _p(dy), _ld(dy), _p(dw), stride_ci, stride_co, N, Do,
                                     Ho, Wo, Cin, Cout, _p(ws), ws.numel(), int(accumulate), _s()), "conv3d_wgrad")


# ----------------------------------------------------------------------------- tcgen05 conv path (bf16)
def tc_supported(mode: int, Cin: int, Cout: int) -> bool:
    return bool(_lib().hdf_tc_supported(mode, Cin, Cout))


def tc_pack(w: torch.Tensor, Cin: int, Cout: int, stride_ci: int, stride_co: int, flip: bool,
            cin_valid: Optional[int] = None) -> torch.Tensor:
    """packed[tap][co][ci] bf16 = w[ci*stride_ci + co*stride_co + (26-tap if flip else tap)], zero for ci >= cin_valid"""
    out = torch.empty((27, Cout, Cin), dtype=torch.bfloat16, device=w.device)
    _C.check(_lib().hdf_tc_pack_weights(_p(w), _p(out), Cin, Cout, stride_ci, stride_co, int(flip),
                                        Cin if cin_valid is None else cin_valid, _s()), "tc_pack")
    return out


def tc_conv3d_fwd(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, mode: int = 0):
    N, Do, Ho, Wo, Cout = out.shape
    Cin = x.shape[-1]
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    _C.check(_lib().hdf_tc_conv3d_fwd(mode, _p(x), _ld(x), _p(wp), _p(bias), _p(out), _ld(out), N, Do, Ho, Wo, Cin, Cout, _s()),
             "tc_conv3d_fwd")
    return out


def tc_convt_supported(Cin: int, Cout: int) -> bool:
    return bool(_lib().hdf_tc_convt_supported(Cin, Cout))


def tc_convt_pack(w: torch.Tensor) -> torch.Tensor:
    """torch ConvTranspose3d weight [64, 32, 3, 3, 3] fp32 -> shared-memory image of the shift-major kernel (bf16)"""
    assert w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape[:2]) == (64, 32)
    out = torch.empty(_lib().hdf_tc_convt_packed_bytes() // 2, dtype=torch.bfloat16, device=w.device)
    _C.check(_lib().hdf_tc_convt_pack_weights(_p(w), _p(out), _s()), "tc_convt_pack")
    return out


def tc_convt_fwd(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor):
    """ConvTranspose3d(64 -> 32, k3, s2, p1, op1): x [N,D,H,W,64] bf16 -> out [N,2D,2H,2W,32] bf16 (may be a channel slice)"""
    N, D, H, W, Cin = x.shape
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and Cin == 64
    assert tuple(out.shape) == (N, 2 * D, 2 * H, 2 * W, 32)
    _C.check(_lib().hdf_tc_convt_fwd(_p(x), _ld(x), _p(wp), _p(bias), _p(out), _ld(out), N, D, H, W, _s()), "tc_convt_fwd")
    return out


def tc_ws_supported(mode: int, Cin: int, Cout: int) -> bool:
    return bool(_lib().hdf_tc_ws_supported(mode, Cin, Cout))


def tc_ws_conv3d_fwd(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor):
    """weight-stationary kernel (Cout = 32, Cin in {32, 64}); same operands as tc_conv3d_fwd(mode 0)"""
    N, D, H, W, Cout = out.shape
    Cin = x.shape[-1]
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and Cout == 32
    _C.check(_lib().hdf_tc_ws_conv3d_fwd(_p(x), _ld(x), _p(wp), _p(bias), _p(out), _ld(out), N, D, H, W, Cin, _s()),
             "tc_ws_conv3d_fwd")
    return out


def tc_ws_conv3d_fwd_stats(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, eps: float = 1e-5):
    """weight-stationary conv + InstanceNorm statistics of its output from the epilogue.  Returns (mean, rstd) [N, 32]."""
    N, D, H, W, Cout = out.shape
    Cin = x.shape[-1]
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and Cout == 32
    mean = torch.empty((N, Cout), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ws = Workspace.get(_lib().hdf_tc_ws_stats_workspace(N))
    _C.check(_lib().hdf_tc_ws_conv3d_fwd_stats(_p(x), _ld(x), _p(wp), _p(bias), _p(out), _ld(out), N, D, H, W, Cin, eps, _p(mean),
                                               _p(rstd), _p(ws), ws.numel(), _s()), "tc_ws_conv3d_fwd_stats")
    return mean, rstd


def tc_wgrad_supported(mode: int, Cin: int, Cout: int) -> bool:
    return bool(_lib().hdf_tc_wgrad_supported(mode, Cin, Cout))


def tc_conv3d_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, stride_ci: int, stride_co: int, mode: int = 0,
                    accumulate=False):
    N, Do, Ho, Wo, Cout = dy.shape
    Cin = x.shape[-1]
    ws = Workspace.get(_lib().hdf_tc_wgrad_workspace(mode, N, Do, Ho, Wo, Cin, Cout))
    _C.check(_lib().hdf_tc_conv3d_wgrad(mode, _p(x), _ld(x), _p(dy), _ld(dy), _p(dw), stride_ci, stride_co, N, Do, Ho, Wo, Cin,
                                        Cout, _p(ws), ws.numel(), int(accumulate), _s()), "tc_conv3d_wgrad")


# ----------------------------------------------------------------------------- stem (first conv through im2col + GEMM)
def stem_kp(Cin: int) -> int:
    return int(_lib().hdf_stem_kp(Cin))


def stem_supported(Cin: int, Cout: int) -> bool:
    return bool(_lib().hdf_stem_supported(Cin, Cout))


def stem_im2col(x: torch.Tensor) -> torch.Tensor:
    """x: NCDHW fp32 -> xcol [N, D, H, W, Kp] bf16 (k = tap*Cin + ci, zero padded)"""
    N, Cin, D, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty((N, D, H, W, stem_kp(Cin)), dtype=torch.bfloat16, device=x.device)
    _C.check(_lib().hdf_stem_im2col(_p(x), _p(out), N, Cin, D, H, W, _s()), "stem_im2col")
    return out


class StemInput:
    """The raw NCDHW fp32 volume standing in for the [N, D, H, W, Kp] bf16 im2col operand of the first convolution: the
    fused kernels (csrc/stem_tc.cu) gather the taps themselves, so no such matrix exists.  Carries the shape / dtype /
    device the engine's bookkeeping reads."""

    def __init__(self, x: torch.Tensor):
        assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 5
        N, Cin, D, H, W = x.shape
        self.raw = x
        self.shape = torch.Size((N, D, H, W, stem_kp(Cin)))
        self.dtype = torch.bfloat16
        self.device = x.device


def stem_fused_supported(Cin: int, Cout: int) -> bool:
    return bool(_lib().hdf_stem_fused_supported(Cin, Cout))


def stem_conv_fwd(xcol, w: torch.Tensor, out: torch.Tensor):
    """w: torch Conv3d weight [Cout, Cin, 3, 3, 3]; out: [N, D, H, W, Cout] bf16 view; xcol: im2col matrix or StemInput"""
    N, D, H, W, Cout = out.shape
    Cin = w.shape[1]
    wp = torch.empty((Cout, xcol.shape[-1]), dtype=torch.bfloat16, device=w.device)
    _C.check(_lib().hdf_stem_pack_weights(_p(w), _p(wp), Cin, Cout, _s()), "stem_pack_weights")
    if isinstance(xcol, StemInput):
        _C.check(_lib().hdf_stem_fused_fwd(_p(xcol.raw), _p(wp), _p(out), _ld(out), N, Cin, D, H, W, Cout, _s()), "stem_fused_fwd")
        return out
    _C.check(_lib().hdf_stem_conv_fwd(_p(xcol), _p(wp), _p(out), _ld(out), N, D, H, W, Cin, Cout, _s()), "stem_conv_fwd")
    return out


def stem_conv_wgrad(xcol, dy: torch.Tensor, dw: torch.Tensor, accumulate=False):
    N, D, H, W, Cout = dy.shape
    Cin = dw.shape[1]
    if isinstance(xcol, StemInput):
        ws = Workspace.get(_lib().hdf_stem_fused_wgrad_workspace(Cin, Cout))
        _C.check(_lib().hdf_stem_fused_wgrad(_p(xcol.raw), _p(dy), _ld(dy), _p(dw), N, Cin, D, H, W, Cout, _p(ws), ws.numel(),
                                             int(accumulate), _s()), "stem_fused_wgrad")
        return
    ws = Workspace.get(_lib().hdf_stem_wgrad_workspace(N, D, H, W, Cin, Cout))
    _C.check(_lib().hdf_stem_conv_wgrad(_p(xcol), _p(dy), _ld(dy), _p(dw), N, D, H, W, Cin, Cout, _p(ws), ws.numel(),
                                        int(accumulate), _s()), "stem_conv_wgrad")


# ----------------------------------------------------------------------------- instance norm & friends
def instnorm_stats(y: torch.Tensor, eps: float = 1e-5):
    N, C = y.shape[0], y.shape[-1]
    V = y.numel() // (N * C)
    mean = torch.empty((N, C), dtype=torch.float32, device=y.device)
    rstd = torch.empty_like(mean)
    ws = Workspace.get(_lib().hdf_reduce_workspace(N, V, C))
    _C.check(_lib().hdf_instnorm_stats(_DT[y.dtype], _p(y), _ld(y), N, V, C, eps, _p(mean), _p(rstd), _p(ws), ws.numel(), _s()),
             "instnorm_stats")
    return mean, rstd


def instnorm_apply(y, mean, rstd, gamma, beta, out, residual=None, relu=True):
    N, C = y.shape[0], y.shape[-1]
    V = y.numel() // (N * C)
    _C.check(_lib().hdf_instnorm_apply(_DT[y.dtype], _p(y), _ld(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(residual),
                                       _ld(residual) if residual is not None else 0, _p(out), _ld(out), N, V, C, int(relu),
                                       _s()), "instnorm_apply")
    return out


def instnorm_apply_head_supported(y: torch.Tensor, ncls: int) -> bool:
    return y.dtype == torch.bfloat16 and bool(_lib().hdf_instnorm_apply_head_supported(y.shape[-1], ncls))


def instnorm_apply_head(y, mean, rstd, gamma, beta, out, head_w, head_b, relu=True):
    """InstanceNorm apply (+affine, ReLU) and the 1x1x1 head on its output in one pass; returns the logits [N, ncls, D, H, W]"""
    N, C = y.shape[0], y.shape[-1]
    V = y.numel() // (N * C)
    ncls = head_w.shape[0]
    logits = torch.empty((N, ncls, *y.shape[1:4]), dtype=y.dtype, device=y.device)
    _C.check(_lib().hdf_instnorm_apply_head(_p(y), _ld(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(out), _ld(out), N, V, C,
                                            int(relu), _p(head_w), _p(head_b), _p(logits), ncls, _s()), "instnorm_apply_head")
    return logits


def instnorm_bwd(dout, y, mean, rstd, gamma, beta, dgamma, dbeta, relu=True, accumulate_params=False):
    N, C = y.shape[0], y.shape[-1]
    V = y.numel() // (N * C)
    dy = torch.empty(y.shape, dtype=y.dtype, device=y.device)
    s1 = torch.empty((N, C), dtype=torch.float32, device=y.device)
    s2 = torch.empty_like(s1)
    ws = Workspace.get(_lib().hdf_reduce_workspace(N, V, C))
    _C.check(_lib().hdf_instnorm_bwd(_DT[y.dtype], _p(dout), _ld(dout), _p(y), _ld(y), _p(mean), _p(rstd), _p(gamma), _p(beta),
                                     _p(dy), _ld(dy), N, V, C, int(relu), _p(s1), _p(s2), _p(dgamma), _p(dbeta),
                                     int(accumulate_params), _p(ws), ws.numel(), _s()), "instnorm_bwd")
    return dy


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate=False):
    C = x.shape[-1]
    rows = x.numel() // C
    ws = Workspace.get(_lib().hdf_reduce_workspace(1, rows, C))
    _C.check(_lib().hdf_colsum(_DT[x.dtype], _p(x), _ld(x), rows, C, _p(out), int(accumulate), _p(ws), ws.numel(), _s()), "colsum")


def add_(dst: torch.Tensor, src: torch.Tensor):
    C = dst.shape[-1]
    _C.check(_lib().hdf_add_(_DT[dst.dtype], _p(dst), _ld(dst), _p(src), _ld(src), dst.numel() // C, C, _s()), "add_")
    return dst


def copy_rows(dst: torch.Tensor, src: torch.Tensor):
    C = dst.shape[-1]
    _C.check(_lib().hdf_copy_rows(_DT[dst.dtype], _p(dst), _ld(dst), _p(src), _ld(src), dst.numel() // C, C, _s()), "copy_rows")
    return dst


def cast_from_f32(src: torch.Tensor, dst: torch.Tensor):
    C = dst.shape[-1]
    _C.check(_lib().hdf_cast_rows_from_f32(_DT[dst.dtype], _p(src), _ld(src), _p(dst), _ld(dst), dst.numel() // C, C, _s()),
             "cast_from_f32")


def cast_to_f32(src: torch.Tensor, dst: torch.Tensor):
    C = dst.shape[-1]
    _C.check(_lib().hdf_cast_rows_to_f32(_DT[src.dtype], _p(src), _ld(src), _p(dst), _ld(dst), dst.numel() // C, C, _s()),
             "cast_to_f32")


def _depth_factor(d_small: int, d_big: int) -> int:
    """2 for the 3-D ops, 1 when the depth axis is left alone (flat [N, 1, H, W, C] volumes of the 2-D model)"""
    f = d_big // max(d_small, 1)
    assert f in (1, 2) and d_small * f == d_big, f"depth {d_big} vs {d_small}"
    return f


def maxpool2_fwd(x: torch.Tensor, out: torch.Tensor):
    N, Do, Ho, Wo, C = out.shape
    pd = _depth_factor(Do, x.shape[1])
    _C.check(_lib().hdf_maxpool2_fwd_ex(_DT[x.dtype], _p(x), _ld(x), _p(out), _ld(out), N, Do, Ho, Wo, C, pd, _s()), "maxpool2_fwd")
    return out


def maxpool2_bwd(x: torch.Tensor, dpool: torch.Tensor, dx: torch.Tensor, accumulate: bool):
    N, Do, Ho, Wo, C = dpool.shape
    pd = _depth_factor(Do, x.shape[1])
    _C.check(_lib().hdf_maxpool2_bwd_ex(_DT[x.dtype], _p(x), _ld(x), _p(dpool), _ld(dpool), _p(dx), _ld(dx), N, Do, Ho, Wo, C,
                                        int(accumulate), pd, _s()), "maxpool2_bwd")
    return dx


def upsample2_fwd(x: torch.Tensor, out: torch.Tensor):
    N, Di, Hi, Wi, C = x.shape
    sd = _depth_factor(Di, out.shape[1])
    _C.check(_lib().hdf_upsample2_fwd_ex(_DT[x.dtype], _p(x), _ld(x), _p(out), _ld(out), N, Di, Hi, Wi, C, sd, _s()), "upsample2_fwd")
    return out


def upsample2_bwd(dout: torch.Tensor, dx: torch.Tensor, accumulate: bool = False):
    N, Di, Hi, Wi, C = dx.shape
    sd = _depth_factor(Di, dout.shape[1])
    _C.check(_lib().hdf_upsample2_bwd_ex(_DT[dx.dtype], _p(dout), _ld(dout), _p(dx), _ld(dx), N, Di, Hi, Wi, C, int(accumulate),
                                         sd, _s()), "upsample2_bwd")
    return dx


def ncdhw_to_cl(x: torch.Tensor, dtype: torch.dtype, pad_to: int = 0) -> torch.Tensor:
    """NCDHW fp32 -> channels-last `dtype`; with pad_to > C the extra channels are zero."""
    N, Cc = x.shape[0], x.shape[1]
    Cp = max(Cc, pad_to)
    alloc = torch.zeros if Cp > Cc else torch.empty
    out = alloc((N, *x.shape[2:], Cp), dtype=dtype, device=x.device)
    V = x.numel() // (N * Cc)
    _C.check(_lib().hdf_ncdhw_to_cl(_DT[dtype], _p(x), _p(out), Cp, N, Cc, V, _s()), "ncdhw_to_cl")
    return out


# ----------------------------------------------------------------------------- heads
def head_fwd(a: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    N, C = a.shape[0], a.shape[-1]
    V = a.numel() // (N * C)
    ncls = w.shape[0]
    out = torch.empty((N, ncls, *a.shape[1:4]), dtype=a.dtype, device=a.device)
    _C.check(_lib().hdf_head_fwd(_DT[a.dtype], _p(a), _ld(a), _p(w), _p(b), _p(out), N, V, C, ncls, _s()), "head_fwd")
    return out


def head_bwd(g: torch.Tensor, a: torch.Tensor, w: torch.Tensor, da: torch.Tensor, dw: torch.Tensor, db: torch.Tensor,
             accumulate_da: bool, accumulate_params: bool = False):
    N, C = a.shape[0], a.shape[-1]
    V = a.numel() // (N * C)
    ncls = w.shape[0]
    assert g.is_contiguous() and g.dtype == a.dtype
    ws = Workspace.get(_lib().hdf_head_bwd_workspace(N, V, C, ncls))
    _C.check(_lib().hdf_head_bwd(_DT[a.dtype], _p(g), _p(a), _ld(a), _p(w), _p(da), _ld(da), _p(dw), _p(db), N, V, C, ncls,
                                 int(accumulate_da), int(accumulate_params), _p(ws), ws.numel(), _s()), "head_bwd")


# ----------------------------------------------------------------------------- tokens (fp32 rows)
def _seed_args(seed):
    """seed: int offset, or a 1-element int64 device tensor holding the base seed (graph-replay safe)"""
    if isinstance(seed, torch.Tensor):
        return _p(seed), 0
    return None, int(seed) & 0xFFFFFFFFFFFFFFFF


def gemm(A, Bm, b_is_nk, out, bias=None, residual=None, pre=None, act=0, p=0.0, seed=0, call_id=0, accumulate=False):
    """out[M,N] = epi(A[M,K] @ (Bm^T if b_is_nk else Bm))"""
    M, K = A.shape
    N = out.shape[1]
    _C.check(_lib().hdf_gemm_rowmajor(_p(A), A.stride(0), _p(Bm), Bm.stride(0), int(b_is_nk), _p(out), out.stride(0), M, N, K,
                                      _p(bias), _p(residual), residual.stride(0) if residual is not None else 0, _p(pre),
                                      act, float(p), *_seed_args(seed), call_id, int(accumulate), _s()), "gemm")
    return out


def gemm_at_b(A, Bm, out, accumulate=True):
    """out[M,N] (+)= A[K,M]^T @ Bm[K,N]   (out dense)"""
    K, M = A.shape
    N = Bm.shape[1]
    assert out.is_contiguous() and out.numel() == M * N
    ws = Workspace.get(_lib().hdf_gemm_at_b_workspace(M, N, K))
    _C.check(_lib().hdf_gemm_at_b(_p(A), A.stride(0), _p(Bm), Bm.stride(0), _p(out), M, N, K, _p(ws), ws.numel(),
                                  int(accumulate), _s()), "gemm_at_b")


def act_dropout_bwd(dy, pre, act, p, seed, call_id):
    M, N = dy.shape
    dz = torch.empty((M, N), dtype=torch.float32, device=dy.device)
    _C.check(_lib().hdf_act_dropout_bwd(_p(dy), dy.stride(0), _p(pre), _p(dz), N, M, N, act, float(p), *_seed_args(seed), call_id, _s()),
             "act_dropout_bwd")
    return dz


def add_rows_f32(dst, src, accumulate):
    rows, Cc = src.shape
    _C.check(_lib().hdf_add_rows_f32(_p(dst), dst.stride(0), _p(src), src.stride(0), rows, Cc, int(accumulate), _s()),
             "add_rows_f32")


def dct_a_fwd(F, Cl, P, q):
    """fused: h0 = F[:, :Cl] Wl^T + bl ; n1 = LN1(h0) ; qkv = n1 Wqkv^T.  Returns (h0, n1, m1, r1, qkv)."""
    R = F.shape[0]
    dev, f32 = F.device, torch.float32
    h0 = torch.empty((R, 32), dtype=f32, device=dev)
    n1 = torch.empty((R, 32), dtype=f32, device=dev)
    m1 = torch.empty((R,), dtype=f32, device=dev)
    r1 = torch.empty((R,), dtype=f32, device=dev)
    qkv = torch.empty((R, 96), dtype=f32, device=dev)
    _C.check(_lib().hdf_dct_a_fwd(_p(F), F.stride(0), Cl, _p(P[q + "0.weight"]), _p(P[q + "0.bias"]), _p(P[q + "1.norm.weight"]),
                                  _p(P[q + "1.norm.bias"]), _p(P[q + "1.fn.to_qkv.weight"]), _p(h0), _p(n1), _p(m1), _p(r1),
                                  _p(qkv), R, _s()), "dct_a_fwd")
    return h0, n1, m1, r1, qkv


def tok_a_fwd(F, Cl, P, q):
    """tensor-core (bf16 path) version of dct_a_fwd: same operands, same saved tensors (csrc/tok_tc.cu)"""
    R = F.shape[0]
    dev, f32 = F.device, torch.float32
    h0 = torch.empty((R, 32), dtype=f32, device=dev)
    n1 = torch.empty((R, 32), dtype=f32, device=dev)
    m1 = torch.empty((R,), dtype=f32, device=dev)
    r1 = torch.empty((R,), dtype=f32, device=dev)
    qkv = torch.empty((R, 96), dtype=f32, device=dev)
    _C.check(_lib().hdf_tok_a_fwd(_p(F), F.stride(0), Cl, _p(P[q + "0.weight"]), _p(P[q + "0.bias"]), _p(P[q + "1.norm.weight"]),
                                  _p(P[q + "1.norm.bias"]), _p(P[q + "1.fn.to_qkv.weight"]), _p(h0), _p(n1), _p(m1), _p(r1),
                                  _p(qkv), R, _s()), "tok_a_fwd")
    return h0, n1, m1, r1, qkv


def tok_c_fwd(qkv, h0, P, q, fout, B, N, scale, p, seed, ids):
    """tensor-core attention (8 heads x 4) fused with the post-attention chain of dct_c_fwd.  Returns (o, lse, saved)."""
    R = qkv.shape[0]
    dev, f32 = qkv.device, torch.float32
    o = torch.empty((R, 32), dtype=f32, device=dev)
    lse = torch.empty((B, 8, N), dtype=f32, device=dev)
    sv = dict(h1=torch.empty((R, 32), dtype=f32, device=dev), n2=torch.empty((R, 32), dtype=f32, device=dev),
              z1=torch.empty((R, 64), dtype=f32, device=dev), f1=torch.empty((R, 64), dtype=f32, device=dev),
              h2=torch.empty((R, 32), dtype=f32, device=dev), n3=torch.empty((R, 32), dtype=f32, device=dev),
              z1b=torch.empty((R, 64), dtype=f32, device=dev), g1=torch.empty((R, 64), dtype=f32, device=dev),
              m2=torch.empty((R,), dtype=f32, device=dev), r2=torch.empty((R,), dtype=f32, device=dev),
              m3=torch.empty((R,), dtype=f32, device=dev), r3=torch.empty((R,), dtype=f32, device=dev))
    sp, so = _seed_args(seed)
    _C.check(_lib().hdf_tok_c_fwd(_p(qkv), _p(h0), _p(o), _p(lse), _p(sv["h1"]), _p(sv["n2"]), _p(sv["z1"]), _p(sv["f1"]), _p(sv["h2"]),
                                  _p(sv["n3"]), _p(sv["z1b"]), _p(sv["g1"]), _p(sv["m2"]), _p(sv["r2"]), _p(sv["m3"]), _p(sv["r3"]),
                                  _p(fout), fout.stride(0), _p(P[q + "1.fn.to_out.0.weight"]), _p(P[q + "1.fn.to_out.0.bias"]),
                                  _p(P[q + "2.norm.weight"]), _p(P[q + "2.norm.bias"]), _p(P[q + "2.fn.net.0.weight"]),
                                  _p(P[q + "2.fn.net.0.bias"]), _p(P[q + "2.fn.net.3.weight"]), _p(P[q + "2.fn.net.3.bias"]), B, N,
                                  float(scale), float(p), sp, so, *ids, _s()), "tok_c_fwd")
    return o, lse, sv


def dct_a_bwd(dqkv, dh1, s, F, Cl, dF, P, G, q):
    """fused backward of dct_a_fwd: dF[:, :Cl] += ..., parameter gradients += into G."""
    R = F.shape[0]
    ws = Workspace.get(_lib().hdf_dct_a_bwd_workspace(R, Cl))
    _C.check(_lib().hdf_dct_a_bwd(_p(dqkv), _p(dh1), _p(s["h0"]), _p(s["n1"]), _p(s["m1"]), _p(s["r1"]), _p(F), F.stride(0), Cl,
                                  _p(P[q + "1.fn.to_qkv.weight"]), _p(P[q + "1.norm.weight"]), _p(P[q + "0.weight"]), _p(dF),
                                  dF.stride(0), _p(G[q + "1.fn.to_qkv.weight"]), _p(G[q + "1.norm.weight"]),
                                  _p(G[q + "1.norm.bias"]), _p(G[q + "0.weight"]), _p(G[q + "0.bias"]), R, _p(ws), ws.numel(), _s()),
             "dct_a_bwd")


def dct_c_fwd(o, h0, P, q, fout, p, seed, ids):
    """fused: h1 = drop(o Wo^T + bo) + h0 ; h2 = FF(LN2(h1)) + h1 ; fout = FF(LN2(h2)).  Returns the saved tensors."""
    R = o.shape[0]
    dev = o.device
    f32 = torch.float32
    sv = dict(h1=torch.empty((R, 32), dtype=f32, device=dev), n2=torch.empty((R, 32), dtype=f32, device=dev),
              z1=torch.empty((R, 64), dtype=f32, device=dev), f1=torch.empty((R, 64), dtype=f32, device=dev),
              h2=torch.empty((R, 32), dtype=f32, device=dev), n3=torch.empty((R, 32), dtype=f32, device=dev),
              z1b=torch.empty((R, 64), dtype=f32, device=dev), g1=torch.empty((R, 64), dtype=f32, device=dev),
              m2=torch.empty((R,), dtype=f32, device=dev), r2=torch.empty((R,), dtype=f32, device=dev),
              m3=torch.empty((R,), dtype=f32, device=dev), r3=torch.empty((R,), dtype=f32, device=dev))
    sp, so = _seed_args(seed)
    _C.check(_lib().hdf_dct_c_fwd(_p(o), _p(h0), _p(sv["h1"]), _p(sv["n2"]), _p(sv["z1"]), _p(sv["f1"]), _p(sv["h2"]), _p(sv["n3"]),
                                  _p(sv["z1b"]), _p(sv["g1"]), _p(sv["m2"]), _p(sv["r2"]), _p(sv["m3"]), _p(sv["r3"]), _p(fout),
                                  fout.stride(0), _p(P[q + "1.fn.to_out.0.weight"]), _p(P[q + "1.fn.to_out.0.bias"]),
                                  _p(P[q + "2.norm.weight"]), _p(P[q + "2.norm.bias"]), _p(P[q + "2.fn.net.0.weight"]),
                                  _p(P[q + "2.fn.net.0.bias"]), _p(P[q + "2.fn.net.3.weight"]), _p(P[q + "2.fn.net.3.bias"]), R,
                                  float(p), sp, so, *ids, _s()), "dct_c_fwd")
    return sv


def dct_c_bwd(dg2, o, sv, P, G, q, p, seed, ids):
    """fused backward of dct_c_fwd: returns (d_o, dh1); parameter gradients are accumulated into G."""
    R = o.shape[0]
    d_o = torch.empty((R, 32), dtype=torch.float32, device=o.device)
    dh1 = torch.empty((R, 32), dtype=torch.float32, device=o.device)
    ws = Workspace.get(_lib().hdf_dct_c_bwd_workspace(R))
    sp, so = _seed_args(seed)
    _C.check(_lib().hdf_dct_c_bwd(_p(dg2), dg2.stride(0), _p(o), _p(sv["h1"]), _p(sv["n2"]), _p(sv["z1"]), _p(sv["f1"]), _p(sv["h2"]),
                                  _p(sv["n3"]), _p(sv["z1b"]), _p(sv["g1"]), _p(sv["m2"]), _p(sv["r2"]), _p(sv["m3"]), _p(sv["r3"]),
                                  _p(P[q + "1.fn.to_out.0.weight"]), _p(P[q + "2.norm.weight"]), _p(P[q + "2.fn.net.0.weight"]),
                                  _p(P[q + "2.fn.net.3.weight"]), _p(d_o), _p(dh1), _p(G[q + "2.fn.net.3.weight"]),
                                  _p(G[q + "2.fn.net.3.bias"]), _p(G[q + "2.fn.net.0.weight"]), _p(G[q + "2.fn.net.0.bias"]),
                                  _p(G[q + "1.fn.to_out.0.weight"]), _p(G[q + "1.fn.to_out.0.bias"]), _p(G[q + "2.norm.weight"]),
                                  _p(G[q + "2.norm.bias"]), R, float(p), sp, so, *ids, _p(ws), ws.numel(), _s()), "dct_c_bwd")
    return d_o, dh1


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    M, Cc = x.shape
    out = torch.empty((M, Cc), dtype=torch.float32, device=x.device)
    mean = torch.empty((M,), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    _C.check(_lib().hdf_layernorm_fwd(_p(x), x.stride(0), _p(gamma), _p(beta), _p(out), Cc, _p(mean), _p(rstd), M, Cc, eps, _s()),
             "layernorm_fwd")
    return out, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, accumulate_dx, dgamma, dbeta):
    M, Cc = x.shape
    ws = Workspace.get(_lib().hdf_layernorm_bwd_workspace(M, Cc))
    _C.check(_lib().hdf_layernorm_bwd(_p(dy), dy.stride(0), _p(x), x.stride(0), _p(mean), _p(rstd), _p(gamma), _p(dx),
                                      dx.stride(0), int(accumulate_dx), _p(dgamma), _p(dbeta), 1, M, Cc, _p(ws), ws.numel(),
                                      _s()), "layernorm_bwd")


def attention_fwd(qkv, B, N, H, scale):
    R, three = qkv.shape
    inner = three // 3
    o = torch.empty((R, inner), dtype=torch.float32, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    _C.check(_lib().hdf_attention_fwd(_p(qkv), qkv.stride(0), _p(o), inner, _p(lse), B, N, H, scale, _s()), "attention_fwd")
    return o, lse


def attention_bwd(qkv, o, dout, lse, B, N, H, scale):
    dqkv = torch.empty_like(qkv)
    _C.check(_lib().hdf_attention_bwd(_p(qkv), qkv.stride(0), _p(o), o.stride(0), _p(dout), dout.stride(0), _p(lse), _p(dqkv),
                                      dqkv.stride(0), B, N, H, scale, _s()), "attention_bwd")
    return dqkv


def patch_embed_fwd(img, modality, weight, bias, pos, out, p, seed, call_id, tensor_cores=False):
    """tokens of one modality: Conv3d(1 -> E, k16, s16) + bias + position embedding, dropout.  tensor_cores=True (bf16
    path): tcgen05 implicit GEMM with bf16 operands (csrc/patch_tc.cu), else the fp32 SIMT GEMM."""
    B, Mch, D, H, W = img.shape
    E = weight.shape[0]
    if tensor_cores and D != 1 and _lib().hdf_patch_embed_tc_supported(E):      # D == 1: 2-D patches (K = 256), SIMT path
        ws = Workspace.get(_lib().hdf_patch_embed_tc_workspace(B, D, H, W, E))
        _C.check(_lib().hdf_patch_embed_tc_fwd(_p(img), B, Mch, modality, D, H, W, _p(weight), _p(bias), _p(pos), _p(out),
                                               out.stride(0), E, float(p), *_seed_args(seed), call_id, _p(ws), ws.numel(), _s()),
                 "patch_embed_tc_fwd")
        return
    ws = Workspace.get(_lib().hdf_patch_embed_fwd_workspace(B, D, H, W, E))
    _C.check(_lib().hdf_patch_embed_fwd(_p(img), B, Mch, modality, D, H, W, _p(weight), _p(bias), _p(pos), _p(out),
                                        out.stride(0), E, float(p), *_seed_args(seed), call_id, _p(ws), ws.numel(), _s()),
             "patch_embed_fwd")


def patch_embed_wgrad(img, modality, dtok, dweight, accumulate=False):
    B, Mch, D, H, W = img.shape
    E = dweight.shape[0]
    ws = Workspace.get(_lib().hdf_patch_embed_wgrad_workspace(B, D, H, W, E))
    _C.check(_lib().hdf_patch_embed_wgrad(_p(img), B, Mch, modality, D, H, W, _p(dtok), dtok.stride(0), _p(dweight), E, _p(ws),
                                          ws.numel(), int(accumulate), _s()), "patch_embed_wgrad")


def posemb_grad(dtok, dpos, B, ntok, E, accumulate=False):
    _C.check(_lib().hdf_posemb_grad(_p(dtok), dtok.stride(0), _p(dpos), B, ntok, E, int(accumulate), _s()), "posemb_grad")


# ----------------------------------------------------------------------------- loss / sliding window
def loss_level_fwd(logits, target, cw, level, ignore_index, smooth, level_weight, ce_w, dice_w, sums, out_level, total):
    B, Cc, Dl, Hl, Wl = logits.shape
    has_ig = ignore_index is not None
    ds = 1 if target.shape[2] == Dl else 1 << level          # flat inputs (2-D model): the depth axis is not strided
    _C.check(_lib().hdf_loss_level_fwd_ex(_DT[logits.dtype], _p(logits), _p(target), _p(cw), B, Cc, Dl, Hl, Wl, 1 << level, ds,
                                          ignore_index if has_ig else -1, int(has_ig), smooth, level_weight, ce_w, dice_w,
                                          _p(sums), _p(out_level), _p(total), _s()), "loss_level_fwd")


def confusion_update(logits: torch.Tensor, target: torch.Tensor, conf: torch.Tensor):
    """conf [B, C, C] int64 += counts of (argmax target, argmax logits) per sample; asynchronous, no host sync"""
    B, Cc = logits.shape[:2]
    V = logits[0, 0].numel()
    assert conf.dtype == torch.int64 and tuple(conf.shape) == (B, Cc, Cc) and conf.is_contiguous()
    _C.check(_lib().hdf_confusion_update(_DT[logits.dtype], _p(logits), _p(target), B, Cc, V, _p(conf), _s()), "confusion_update")
    return conf


def loss_level_bwd(logits, target, cw, level, ignore_index, smooth, level_weight, ce_w, dice_w, sums, grad_out, dlogits):
    B, Cc, Dl, Hl, Wl = logits.shape
    has_ig = ignore_index is not None
    ds = 1 if target.shape[2] == Dl else 1 << level
    _C.check(_lib().hdf_loss_level_bwd_ex(_DT[logits.dtype], _p(logits), _p(target), _p(cw), B, Cc, Dl, Hl, Wl, 1 << level, ds,
                                          ignore_index if has_ig else -1, int(has_ig), smooth, level_weight, ce_w, dice_w,
                                          _p(sums), _p(grad_out), _p(dlogits), _s()), "loss_level_bwd")


def sw_accumulate(logits, agg, x0, y0, z0):
    _, Cc, px, py, pz = logits.shape
    _, X, Y, Z = agg.shape
    _C.check(_lib().hdf_sw_accumulate(_DT[logits.dtype], _p(logits), _p(agg), Cc, X, Y, Z, x0, y0, z0, px, py, pz, _s()),
             "sw_accumulate")


def sw_finalize(agg, steps, patch, normalise=True):
    Cc, X, Y, Z = agg.shape
    mask = torch.empty((X, Y, Z), dtype=torch.int64, device=agg.device)
    arr = [(ctypes.c_int * len(s))(*s) for s in steps]
    _C.check(_lib().hdf_sw_finalize(_p(agg), _p(mask), Cc, X, Y, Z, arr[0], len(steps[0]), arr[1], len(steps[1]), arr[2],
                                    len(steps[2]), patch[0], patch[1], patch[2], int(normalise), _s()), "sw_finalize")
    return mask


# ----------------------------------------------------------------------------- input pipeline (csrc/prep.cu)
NORM_MODES = {None: 0, "none": 0, "petct": 1, "mr": 2, "trunc": 3}


def prep_sample(vol: torch.Tensor, lab: Optional[torch.Tensor], origin, size, norm: Optional[str], p0: float, p1: float,
                affine: Optional[torch.Tensor], flip_axis: int, num_class: int, img_out: torch.Tensor,
                lab_out: Optional[torch.Tensor]):
    """One sample of the fused crop -> normalise -> warp -> flip -> one-hot chain.  vol [M, Dv, Hv, Wv] fp32, lab
    [Dv, Hv, Wv] fp32 (device, contiguous); affine: device double [3, 4] or None; outputs [M, D, H, W] / [C, D, H, W] fp32."""
    ensure_init(vol)
    M, Dv, Hv, Wv = vol.shape
    D, H, W = size
    assert vol.dtype == torch.float32 and vol.is_contiguous() and img_out.dtype == torch.float32 and img_out.is_contiguous()
    assert tuple(img_out.shape) == (M, D, H, W)
    if lab is not None:
        assert lab.dtype == torch.float32 and lab.is_contiguous() and tuple(lab.shape) == (Dv, Hv, Wv)
    if lab_out is not None:
        assert lab_out.dtype == torch.float32 and lab_out.is_contiguous() and tuple(lab_out.shape) == (num_class, D, H, W)
    if affine is not None:
        assert affine.dtype == torch.float64 and affine.is_cuda and affine.is_contiguous() and affine.numel() == 12
    ws = Workspace.get(_lib().hdf_prep_workspace(M))
    _C.check(_lib().hdf_prep_sample(_p(vol), _p(lab), M, Dv, Hv, Wv, int(origin[0]), int(origin[1]), int(origin[2]), D, H, W,
                                    NORM_MODES[norm], float(p0), float(p1), _p(affine), int(flip_axis), int(num_class),
                                    _p(img_out), _p(lab_out), _p(ws), ws.numel(), _s()), "prep_sample")
    return img_out, lab_out
