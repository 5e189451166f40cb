"""Per-kernel parity on the GPU, through the C ABI (hdenseformer_b200.ops -> libhdf_b200.so), against plain
torch fp32 ops of the same semantics (the reference's building blocks).  fp32 path tolerance 1e-4 relative
(north_star), bf16 storage path 2e-2."""
import math

import numpy as np

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hdenseformer_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def to_cl(x, dt):  # NCDHW -> NDHWC
    return x.permute(0, 2, 3, 4, 1).contiguous().to(dt)


def from_cl(x):
    return x.float().permute(0, 4, 1, 2, 3).contiguous()


@pytest.fixture(autouse=True)
def _init():
    ops.ensure_init(torch.zeros(1, device=DEV))
    torch.manual_seed(0)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cin,cout,size", [(2, 16, (8, 12, 10)), (16, 32, (9, 9, 9)), (32, 16, (6, 10, 18)), (3, 8, (5, 7, 6))])
def test_conv3d_k3_fwd_dgrad_wgrad(dt, cin, cout, size):
    x = torch.randn(2, cin, *size, device=DEV)
    w = torch.randn(cout, cin, 3, 3, 3, device=DEV) / math.sqrt(27 * cin)
    b = torch.randn(cout, device=DEV)
    xq = to_cl(x, dt)
    xr = from_cl(xq).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv3d(xr, wr, b, padding=1)
    wp = ops.conv_pack(w, cin, cout, 27, cin * 27, False)
    y = torch.empty((2, *size, cout), dtype=dt, device=DEV)
    ops.conv3d_fwd(xq, wp, b, y, 0)
    assert rel(from_cl(y), ref) < TOL[dt]
    g = torch.randn_like(ref)
    gq = to_cl(g, dt)
    ref.backward(from_cl(gq))
    wpd = ops.conv_pack(w, cout, cin, cin * 27, 27, True)
    dx = torch.empty((2, *size, cin), dtype=dt, device=DEV)
    ops.conv3d_fwd(gq, wpd, None, dx, 0)
    assert rel(from_cl(dx), xr.grad) < TOL[dt]
    dw = torch.zeros_like(w)
    ops.conv3d_wgrad(xq, gq, dw, 27, cin * 27, 0)
    assert rel(dw, wr.grad) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_conv_transpose_fwd_dgrad_wgrad(dt):
    cin, cout, size = 16, 8, (4, 6, 5)
    x = torch.randn(2, cin, *size, device=DEV)
    w = torch.randn(cin, cout, 3, 3, 3, device=DEV) / math.sqrt(8 * cin)
    b = torch.randn(cout, device=DEV)
    xq = to_cl(x, dt)
    xr = from_cl(xq).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv_transpose3d(xr, wr, b, stride=2, padding=1, output_padding=1)
    osz = tuple(2 * s for s in size)
    wp = ops.conv_pack(w, cin, cout, cout * 27, 27, False)
    buf = torch.zeros((2, *osz, cout + 8), dtype=dt, device=DEV)      # write into a channel slice
    y = buf[..., 8:]
    ops.conv3d_fwd(xq, wp, b, y, 1)
    assert rel(from_cl(y), ref) < TOL[dt]
    assert buf[..., :8].abs().max().item() == 0
    g = torch.randn_like(ref)
    gq = to_cl(g, dt)
    ref.backward(from_cl(gq))
    wpd = ops.conv_pack(w, cout, cin, 27, cout * 27, False)
    dx = torch.empty((2, *size, cin), dtype=dt, device=DEV)
    ops.conv3d_fwd(gq, wpd, None, dx, 2)
    assert rel(from_cl(dx), xr.grad) < TOL[dt]
    dw = torch.zeros_like(w)
    ops.conv3d_wgrad(xq, gq, dw, cout * 27, 27, 1)
    assert rel(dw, wr.grad) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("affine,res", [(True, True), (False, False)])
def test_instnorm_relu_fwd_bwd(dt, affine, res):
    C, size = 16, (6, 5, 7)
    y = torch.randn(2, C, *size, device=DEV) * 2 + 0.5
    r = torch.randn(2, C, *size, device=DEV)
    gm = (1 + 0.1 * torch.randn(C, device=DEV)) if affine else None
    bt = (0.1 * torch.randn(C, device=DEV)) if affine else None
    yq, rq = to_cl(y, dt), to_cl(r, dt)
    yr = from_cl(yq).requires_grad_(True)
    gr = gm.clone().requires_grad_(True) if affine else None
    br = bt.clone().requires_grad_(True) if affine else None
    ref = F.relu(F.instance_norm(yr, weight=gr, bias=br, eps=1e-5))
    if res:
        ref = ref + from_cl(rq)
    mean, rstd = ops.instnorm_stats(yq)
    out = torch.empty_like(yq)
    ops.instnorm_apply(yq, mean, rstd, gm, bt, out, residual=rq if res else None, relu=True)
    assert rel(from_cl(out), ref) < TOL[dt]
    g = torch.randn_like(ref)
    gq = to_cl(g, dt)
    ref.backward(from_cl(gq))
    dg = torch.zeros(C, device=DEV) if affine else None
    db = torch.zeros(C, device=DEV) if affine else None
    dy = ops.instnorm_bwd(gq, yq, mean, rstd, gm, bt, dg, db, relu=True)
    assert rel(from_cl(dy), yr.grad) < (1e-3 if dt == torch.float32 else 3e-2)
    if affine:
        assert rel(dg, gr.grad) < TOL[dt] * 5 and rel(db, br.grad) < TOL[dt] * 5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_maxpool_and_upsample(dt):
    C, size = 8, (4, 6, 8)
    x = torch.randn(2, C, *size, device=DEV)
    x[:, :, :2] = 0.0   # ties: first max in scan order must win, like torch
    xq = to_cl(x, dt)
    xr = from_cl(xq).requires_grad_(True)
    ref = F.max_pool3d(xr, 2, 2)
    out = torch.empty((2, 2, 3, 4, C), dtype=dt, device=DEV)
    ops.maxpool2_fwd(xq, out)
    assert torch.equal(from_cl(out), ref.detach())
    g = to_cl(torch.randn_like(ref), dt)
    ref.backward(from_cl(g))
    dx = torch.full_like(xq, 7.0)
    ops.maxpool2_bwd(xq, g, dx, False)
    assert torch.equal(from_cl(dx), xr.grad)
    base = to_cl(torch.randn_like(x), dt)
    dx2 = base.clone()
    ops.maxpool2_bwd(xq, g, dx2, True)
    assert rel(from_cl(dx2), from_cl(base) + xr.grad) < TOL[dt]
    # trilinear x2
    xr2 = from_cl(xq).requires_grad_(True)
    ref = F.interpolate(xr2, scale_factor=2, mode="trilinear", align_corners=False)
    up = torch.empty((2, 8, 12, 16, C), dtype=dt, device=DEV)
    ops.upsample2_fwd(xq, up)
    assert rel(from_cl(up), ref) < TOL[dt]
    g = to_cl(torch.randn_like(ref), dt)
    ref.backward(from_cl(g))
    dxu = torch.empty_like(xq)
    ops.upsample2_bwd(g, dxu, False)
    assert rel(from_cl(dxu), xr2.grad) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,ncls", [(16, 2), (32, 4), (256, 3)])
def test_head_fwd_bwd(dt, C, ncls):
    size = (4, 5, 6)
    a = torch.randn(2, C, *size, device=DEV)
    w = torch.randn(ncls, C, 1, 1, 1, device=DEV) / math.sqrt(C)
    b = torch.randn(ncls, device=DEV)
    aq = to_cl(a, dt)
    ar = from_cl(aq).requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv3d(ar, wr, br)
    out = ops.head_fwd(aq, w.view(ncls, C), b)
    assert out.shape == ref.shape and rel(out, ref) < TOL[dt]
    g = torch.randn_like(ref).to(dt)
    ref.backward(g.float())
    base = to_cl(torch.randn_like(a), dt)
    da = base.clone()
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.head_bwd(g.contiguous(), aq, w.view(ncls, C), da, dw, db, True)
    assert rel(from_cl(da), from_cl(base) + ar.grad) < TOL[dt]
    assert rel(dw, wr.grad) < TOL[dt] and rel(db, br.grad) < TOL[dt]


def test_token_gemm_layernorm_attention():
    R, B = 2 * 27, 2
    x = torch.randn(R, 160, device=DEV)
    W = torch.randn(32, 96, device=DEV) / 10
    b = torch.randn(32, device=DEV)
    out = torch.empty(R, 32, device=DEV)
    pre = torch.empty(R, 32, device=DEV)
    res = torch.randn(R, 32, device=DEV)
    ops.gemm(x[:, :96], W, True, out, bias=b, residual=res, pre=pre, act=1)
    z = x[:, :96] @ W.t() + b
    assert rel(pre, z) < 1e-5 and rel(out, F.gelu(z) + res) < 1e-5
    dW = torch.zeros(32, 96, device=DEV)
    g = torch.randn(R, 32, device=DEV)
    ops.gemm_at_b(g, x[:, :96], dW, accumulate=False)
    assert rel(dW, g.t() @ x[:, :96]) < 1e-5
    dx = torch.zeros(R, 160, device=DEV)
    ops.gemm(g, W, False, dx[:, :96])
    assert rel(dx[:, :96], g @ W) < 1e-5 and dx[:, 96:].abs().max().item() == 0
    # gelu backward
    dz = ops.act_dropout_bwd(g, pre, 1, 0.0, 0, 0)
    zr = z.clone().requires_grad_(True)
    F.gelu(zr).backward(g)
    assert rel(dz, zr.grad) < 1e-5
    # layer norm
    h = torch.randn(R, 32, device=DEV, requires_grad=True)
    gm, bt = torch.randn(32, device=DEV, requires_grad=True), torch.randn(32, device=DEV, requires_grad=True)
    ref = F.layer_norm(h, (32,), gm, bt, 1e-5)
    o, m, r = ops.layernorm_fwd(h.detach(), gm.detach(), bt.detach())
    assert rel(o, ref) < 1e-5
    ref.backward(g)
    dh = torch.empty(R, 32, device=DEV)
    dgm, dbt = torch.zeros(32, device=DEV), torch.zeros(32, device=DEV)
    ops.layernorm_bwd(g, h.detach(), m, r, gm.detach(), dh, False, dgm, dbt)
    assert rel(dh, h.grad) < 1e-4 and rel(dgm, gm.grad) < 1e-4 and rel(dbt, bt.grad) < 1e-4
    # attention, 8 heads of dim 4
    for N in (27, 200):
        qkv = torch.randn(B * N, 96, device=DEV, requires_grad=True)
        q, k, v = (t.reshape(B, N, 8, 4).permute(0, 2, 1, 3) for t in qkv.chunk(3, -1))
        att = ((q @ k.transpose(-1, -2)) * 0.5).softmax(-1)
        ref = (att @ v).permute(0, 2, 1, 3).reshape(B * N, 32)
        o, lse = ops.attention_fwd(qkv.detach(), B, N, 8, 0.5)
        assert rel(o, ref) < 1e-5
        go = torch.randn_like(ref)
        ref.backward(go)
        dqkv = ops.attention_bwd(qkv.detach(), o, go, lse, B, N, 8, 0.5)
        assert rel(dqkv, qkv.grad) < 1e-4


def test_dropout_statistics_and_mask_replay():
    R, N = 4096, 64
    a = torch.ones(R, 8, device=DEV)
    W = torch.ones(N, 8, device=DEV)
    outs = []
    for cid in (1, 2):
        out = torch.empty(R, N, device=DEV)
        ops.gemm(a, W, True, out, p=0.5, seed=1234, call_id=cid)
        outs.append(out)
        keep = (out != 0).float().mean().item()
        assert abs(keep - 0.5) < 0.01                      # keep rate
        assert torch.all((out == 0) | (out == 16.0))       # inverted-dropout scale 2
    assert 0.45 < ((outs[0] != 0) == (outs[1] != 0)).float().mean().item() < 0.55   # call sites independent
    g = torch.ones(R, N, device=DEV)
    dz = ops.act_dropout_bwd(g, None, 0, 0.5, 1234, 1)
    assert torch.equal(dz != 0, outs[0] != 0)              # backward replays the forward mask


def test_patch_embed_fwd_wgrad():
    B, Mch, size, E = 2, 2, (32, 16, 48), 32
    img = torch.randn(B, Mch, *size, device=DEV)
    w = (torch.randn(E, 1, 16, 16, 16, device=DEV) / 64).requires_grad_(True)
    b = torch.randn(E, device=DEV)
    ntok = 2 * 1 * 3
    pos = torch.randn(1, ntok, E, device=DEV)
    out = torch.zeros(B * ntok, E + 16, device=DEV)
    ops.patch_embed_fwd(img, 1, w.detach(), b, pos, out[:, :E], 0.0, 0, 0)
    ref = F.conv3d(img[:, 1:2], w, b, stride=16).flatten(2).transpose(1, 2) + pos
    assert rel(out[:, :E], ref.reshape(B * ntok, E)) < 1e-5
    g = torch.randn(B * ntok, E, device=DEV)
    ref.reshape(B * ntok, E).backward(g)
    dw = torch.zeros_like(w)
    ops.patch_embed_wgrad(img, 1, g, dw)
    assert rel(dw, w.grad) < 1e-5
    dpos = torch.zeros_like(pos)
    ops.posemb_grad(g, dpos, B, ntok, E)
    assert rel(dpos, g.view(B, ntok, E).sum(0, keepdim=True)) < 1e-5


@pytest.mark.parametrize("B,size,E", [(2, (144, 144, 144), 128), (1, (32, 16, 48), 64), (3, (16, 48, 32), 128), (1, (96, 96, 96), 256)])
def test_patch_embed_tcgen05_matches_torch_and_simt(B, size, E):
    """csrc/patch_tc.cu (tcgen05 implicit GEMM: producer-gathered A tiles, TMA weights, split-K 8) against F.conv3d in fp32 on
    bf16-rounded operands (products exact, only the fp32 accumulation order differs: 1e-5) and against the fp32 SIMT kernel
    with dropout on: the same counter-based mask must be applied (zeros in the same places)."""
    torch.manual_seed(B * 7 + E)
    Mch = 2
    img = torch.randn(B, Mch, *size, device=DEV)
    w = torch.randn(E, 1, 16, 16, 16, device=DEV) / 64
    b = torch.randn(E, device=DEV)
    ntok = (size[0] // 16) * (size[1] // 16) * (size[2] // 16)
    pos = torch.randn(1, ntok, E, device=DEV)
    out = torch.zeros(B * ntok, E + 16, device=DEV)
    ops.patch_embed_fwd(img, 1, w, b, pos, out[:, :E], 0.0, 0, 0, tensor_cores=True)
    ref = F.conv3d(img[:, 1:2].bfloat16().float(), w.bfloat16().float(), b, stride=16).flatten(2).transpose(1, 2) + pos
    assert rel(out[:, :E], ref.reshape(B * ntok, E)) < 1e-5
    assert out[:, E:].abs().max().item() == 0
    seed = torch.tensor([4242], dtype=torch.int64, device=DEV)
    o_tc, o_simt = torch.empty(B * ntok, E, device=DEV), torch.empty(B * ntok, E, device=DEV)
    ops.patch_embed_fwd(img, 0, w, b, pos, o_tc, 0.5, seed, 9, tensor_cores=True)
    ops.patch_embed_fwd(img, 0, w, b, pos, o_simt, 0.5, seed, 9, tensor_cores=False)
    assert torch.equal(o_tc == 0, o_simt == 0)
    assert rel(o_tc, o_simt) < 1e-2                       # bf16 operand rounding


def test_fused_dct_chain_matches_unfused_composition():
    """hdf_dct_c_fwd / hdf_dct_c_bwd (one kernel each) against the same chain composed of single-op kernels, with
    dropout ON (same counter-based masks on both sides) and a ragged row count."""
    from hdenseformer_b200.engine import Config, Engine
    torch.manual_seed(5)
    R, E = 2 * 27 + 5, 64
    q = "l."
    names = {"1.fn.to_out.0.weight": (32, 32), "1.fn.to_out.0.bias": (32,), "2.norm.weight": (32,), "2.norm.bias": (32,),
             "2.fn.net.0.weight": (64, 32), "2.fn.net.0.bias": (64,), "2.fn.net.3.weight": (32, 64), "2.fn.net.3.bias": (32,)}
    P = {q + k: (torch.randn(s, device=DEV) * (0.2 if len(s) > 1 else 0.1) + (1.0 if k.endswith("norm.weight") else 0.0))
         for k, s in names.items()}
    o, h0 = torch.randn(R, 32, device=DEV), torch.randn(R, 32, device=DEV)
    ids = (3, 4, 5, 6, 7)
    seed = torch.tensor([987654321], dtype=torch.int64, device=DEV)
    F1 = torch.zeros(R, 96, device=DEV)
    sv = ops.dct_c_fwd(o, h0, P, q, F1[:, 32:64], 0.5, seed, ids)
    # unfused forward with the same ids
    pdrop = 0.5
    h1 = torch.empty(R, 32, device=DEV)
    ops.gemm(o, P[q + "1.fn.to_out.0.weight"], True, h1, bias=P[q + "1.fn.to_out.0.bias"], residual=h0, p=pdrop, seed=seed, call_id=ids[0])
    n2, m2, r2 = ops.layernorm_fwd(h1, P[q + "2.norm.weight"], P[q + "2.norm.bias"])
    z1, f1 = torch.empty(R, 64, device=DEV), torch.empty(R, 64, device=DEV)
    ops.gemm(n2, P[q + "2.fn.net.0.weight"], True, f1, bias=P[q + "2.fn.net.0.bias"], pre=z1, act=1, p=pdrop, seed=seed, call_id=ids[1])
    h2 = torch.empty(R, 32, device=DEV)
    ops.gemm(f1, P[q + "2.fn.net.3.weight"], True, h2, bias=P[q + "2.fn.net.3.bias"], residual=h1, p=pdrop, seed=seed, call_id=ids[2])
    n3, m3, r3 = ops.layernorm_fwd(h2, P[q + "2.norm.weight"], P[q + "2.norm.bias"])
    z1b, g1 = torch.empty(R, 64, device=DEV), torch.empty(R, 64, device=DEV)
    ops.gemm(n3, P[q + "2.fn.net.0.weight"], True, g1, bias=P[q + "2.fn.net.0.bias"], pre=z1b, act=1, p=pdrop, seed=seed, call_id=ids[3])
    F2 = torch.zeros(R, 96, device=DEV)
    ops.gemm(g1, P[q + "2.fn.net.3.weight"], True, F2[:, 32:64], bias=P[q + "2.fn.net.3.bias"], p=pdrop, seed=seed, call_id=ids[4])
    for a, b in ((sv["h1"], h1), (sv["n2"], n2), (sv["z1"], z1), (sv["f1"], f1), (sv["h2"], h2), (sv["n3"], n3), (sv["z1b"], z1b),
                 (sv["g1"], g1), (F1, F2)):
        assert rel(a, b) < 1e-5
    assert (F1[:, :32].abs().max() + F1[:, 64:].abs().max()).item() == 0
    # backward
    dF = torch.randn(R, 96, device=DEV)
    Gf = {k: torch.zeros_like(v) for k, v in P.items()}
    Gu = {k: torch.zeros_like(v) for k, v in P.items()}
    do_f, dh_f = ops.dct_c_bwd(dF[:, 32:64], o, sv, P, Gf, q, pdrop, seed, ids)
    eng = Engine(Config(2, 2, 16, (32, 32, 32), 4))
    s = dict(o=o, h1=h1, n2=n2, m2=m2, r2=r2, z1=z1, f1=f1, h2=h2, n3=n3, m3=m3, r3=r3, z1b=z1b, g1=g1, ids=ids)
    do_u, dh_u = eng._dct_c_bwd_unfused(P, Gu, q, s, dF[:, 32:64], R, pdrop, seed)
    assert rel(do_f, do_u) < 1e-4 and rel(dh_f, dh_u) < 1e-4
    for k in P:
        assert rel(Gf[k], Gu[k]) < 1e-4, k


def test_fused_dct_head_matches_unfused_composition():
    """hdf_dct_a_fwd / hdf_dct_a_bwd against gemm + layernorm + gemm composed from single-op kernels."""
    torch.manual_seed(6)
    R, FW = 2 * 27 + 3, 256
    for Cl in (128, 224):
        q = "l."
        P = {q + "0.weight": torch.randn(32, Cl, device=DEV) * 0.1, q + "0.bias": torch.randn(32, device=DEV) * 0.1,
             q + "1.norm.weight": 1 + 0.1 * torch.randn(32, device=DEV), q + "1.norm.bias": 0.1 * torch.randn(32, device=DEV),
             q + "1.fn.to_qkv.weight": torch.randn(96, 32, device=DEV) * 0.2}
        F = torch.randn(R, FW, device=DEV)
        h0, n1, m1, r1, qkv = ops.dct_a_fwd(F, Cl, P, q)
        h0u = torch.empty(R, 32, device=DEV)
        ops.gemm(F[:, :Cl], P[q + "0.weight"], True, h0u, bias=P[q + "0.bias"])
        n1u, m1u, r1u = ops.layernorm_fwd(h0u, P[q + "1.norm.weight"], P[q + "1.norm.bias"])
        qkvu = torch.empty(R, 96, device=DEV)
        ops.gemm(n1u, P[q + "1.fn.to_qkv.weight"], True, qkvu)
        assert rel(h0, h0u) < 1e-5 and rel(n1, n1u) < 1e-5 and rel(qkv, qkvu) < 1e-5 and rel(r1, r1u) < 1e-5
        dqkv, dh1 = torch.randn(R, 96, device=DEV), torch.randn(R, 32, device=DEV)
        base = torch.randn(R, FW, device=DEV)
        dF_f, dF_u = base.clone(), base.clone()
        Gf = {k: torch.zeros_like(v) for k, v in P.items()}
        Gu = {k: torch.zeros_like(v) for k, v in P.items()}
        ops.dct_a_bwd(dqkv, dh1, dict(h0=h0, n1=n1, m1=m1, r1=r1), F, Cl, dF_f, P, Gf, q)
        dh = dh1.clone()
        ops.gemm_at_b(dqkv, n1u, Gu[q + "1.fn.to_qkv.weight"])
        dn1 = torch.empty(R, 32, device=DEV)
        ops.gemm(dqkv, P[q + "1.fn.to_qkv.weight"], False, dn1)
        ops.layernorm_bwd(dn1, h0u, m1u, r1u, P[q + "1.norm.weight"], dh, True, Gu[q + "1.norm.weight"], Gu[q + "1.norm.bias"])
        ops.gemm_at_b(dh, F[:, :Cl], Gu[q + "0.weight"])
        ops.colsum(dh, Gu[q + "0.bias"], accumulate=True)
        ops.gemm(dh, P[q + "0.weight"], False, dF_u[:, :Cl], accumulate=True)
        assert rel(dF_f, dF_u) < 1e-4
        assert torch.equal(dF_f[:, Cl:], base[:, Cl:])
        for k in P:
            assert rel(Gf[k], Gu[k]) < 1e-4, k


def test_upsample_cell_kernel_odd_sizes_and_channel_slices():
    """bf16 cell-centred trilinear x2 (csrc/glue.cu): odd spatial sizes, size-1 dims (all-clamped cells), input and output
    as channel slices of wider buffers."""
    torch.manual_seed(11)
    for size, C in [((3, 5, 7), 32), ((1, 1, 9), 16), ((2, 9, 1), 64)]:
        x = torch.randn(2, *size, C + 16, device=DEV).to(torch.bfloat16)
        xs = x[..., 8:8 + C]
        ref = F.interpolate(xs.float().permute(0, 4, 1, 2, 3), scale_factor=2, mode="trilinear", align_corners=False)
        buf = torch.full((2, 2 * size[0], 2 * size[1], 2 * size[2], C + 8), 5.0, dtype=torch.bfloat16, device=DEV)
        ops.upsample2_fwd(xs, buf[..., :C])
        got = buf[..., :C].float().permute(0, 4, 1, 2, 3)
        assert rel(got, ref) < TOL[torch.bfloat16]
        assert (buf[..., C:] == 5.0).all()


def test_upsample_bwd_odd_sizes_slices_and_accumulate():
    """bf16 trilinear x2 backward (csrc/glue.cu; gather form by default, separable form with HDF_UPS_BWD_V2=1) vs torch
    autograd: odd sizes, size-1 dims, channel-slice operands, overwrite and accumulate modes."""
    torch.manual_seed(12)
    for size, C in [((3, 5, 7), 32), ((1, 1, 9), 16), ((2, 9, 1), 64), ((4, 6, 8), 8)]:
        x = torch.randn(2, C, *size, device=DEV, requires_grad=True)
        up = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        gfull = torch.randn(2, 2 * size[0], 2 * size[1], 2 * size[2], C + 8, device=DEV).to(torch.bfloat16)
        g = gfull[..., 4:4 + C]                                    # channel slice (ld != C)
        up.backward(g.float().permute(0, 4, 1, 2, 3))
        ref = x.grad.permute(0, 2, 3, 4, 1)
        buf = torch.full((2, *size, C + 16), 3.0, dtype=torch.bfloat16, device=DEV)
        ops.upsample2_bwd(g, buf[..., 8:8 + C], False)
        assert rel(buf[..., 8:8 + C].float(), ref) < TOL[torch.bfloat16]
        assert (buf[..., :8] == 3.0).all() and (buf[..., 8 + C:] == 3.0).all()
        base = torch.randn(2, *size, C, device=DEV).to(torch.bfloat16)
        acc = base.clone()
        ops.upsample2_bwd(g, acc, True)
        assert rel(acc.float(), base.float() + ref) < 2 * TOL[torch.bfloat16]


@pytest.mark.parametrize("B,N,Cl,pdrop", [(2, 729, 128, 0.0), (1, 216, 224, 0.5), (2, 37, 160, 0.5)])
def test_token_tensor_core_layer_matches_simt_kernels(B, N, Cl, pdrop):
    """bf16-path tensor-core forward of one DCT inner layer (csrc/tok_tc.cu: mma.sync bf16 Linears, tf32 Q K^T / P V, online
    softmax in registers) against the fp32 SIMT kernels on the same operands and the same dropout masks.  Tolerance = bf16
    operand rounding (2^-9 relative per product, fp32 accumulation)."""
    torch.manual_seed(B * 1000 + N)
    R = B * N
    q = "l."
    names = {"0.weight": (32, Cl), "0.bias": (32,), "1.norm.weight": (32,), "1.norm.bias": (32,), "1.fn.to_qkv.weight": (96, 32),
             "1.fn.to_out.0.weight": (32, 32), "1.fn.to_out.0.bias": (32,), "2.norm.weight": (32,), "2.norm.bias": (32,),
             "2.fn.net.0.weight": (64, 32), "2.fn.net.0.bias": (64,), "2.fn.net.3.weight": (32, 64), "2.fn.net.3.bias": (32,)}
    P = {q + k: (torch.randn(s, device=DEV) * (1.5 / (s[-1] ** 0.5) if len(s) > 1 else 0.1) + (1.0 if k.endswith("norm.weight") else 0.0))
         for k, s in names.items()}
    F1 = torch.randn(R, 256, device=DEV)
    F2 = F1.clone()
    ids = (3, 4, 5, 6, 7)
    seed = torch.tensor([13579], dtype=torch.int64, device=DEV)
    scale = 4 ** -0.5
    # reference: fp32 SIMT kernels
    h0, n1, m1, r1, qkv = ops.dct_a_fwd(F1, Cl, P, q)
    o, lse = ops.attention_fwd(qkv, B, N, 8, scale)
    sv = ops.dct_c_fwd(o, h0, P, q, F1[:, Cl:Cl + 32], pdrop, seed, ids)
    # tensor-core kernels
    h0t, n1t, m1t, r1t, qkvt = ops.tok_a_fwd(F2, Cl, P, q)
    for a, b, tol in ((h0t, h0, 1e-2), (n1t, n1, 1.5e-2), (qkvt, qkv, 2e-2), (m1t, m1, 2e-2)):
        assert rel(a, b) < tol, rel(a, b)
    # attention + chain fed with the SAME qkv / h0 so that only this kernel's arithmetic differs
    ot, lset, svt = ops.tok_c_fwd(qkv, h0, P, q, F2[:, Cl:Cl + 32], B, N, scale, pdrop, seed, ids)
    torch.cuda.synchronize()
    assert rel(ot, o) < 5e-3, rel(ot, o)                    # tf32 scores / probabilities, fp32 statistics
    assert (lset - lse).abs().max().item() < 2e-2           # tf32 scores: |s| ~ 10 x 2^-11
    for k in ("h1", "n2", "z1", "f1", "h2", "n3", "z1b", "g1"):
        assert rel(svt[k], sv[k]) < 3e-2, (k, rel(svt[k], sv[k]))
    assert rel(F2[:, Cl:Cl + 32], F1[:, Cl:Cl + 32]) < 3e-2
    assert torch.equal(F2[:, :Cl], F1[:, :Cl]) and torch.equal(F2[:, Cl + 32:], F1[:, Cl + 32:])
    if pdrop > 0:     # identical masks: the dropped positions coincide (up to gelu underflow at z < -5.9)
        assert ((svt["f1"] == 0) != (sv["f1"] == 0)).float().mean().item() < 1e-3
        assert abs((svt["f1"] == 0).float().mean().item() - pdrop) < 0.03


@pytest.mark.parametrize("C,B,size,dt", [(2, 2, (16, 12, 20), torch.float32), (4, 3, (8, 9, 10), torch.bfloat16), (3, 1, (5, 7, 11), torch.float32)])
def test_on_device_metric_tail_matches_reference_semantics(C, B, size, dt):
    """hdf_confusion_update + metrics.compute_dice / RunningDice (no host sync per step) against the restated reference
    (oracle.compute_dice = trainer.py:919-945; numpy confusion matrix = metrics.py:82-151), incl. a class absent from both
    masks and a sample whose ground truth is background only."""
    from oracle import hdf_oracle as O
    from hdenseformer_b200 import metrics as M
    torch.manual_seed(C * 10 + B)
    logits = torch.randn(B, C, *size, device=DEV)
    if C > 2:
        logits[:, C - 1] -= 50.0                      # last class never predicted ...
    lab = torch.randint(0, C - 1 if C > 2 else C, (B, *size), device=DEV)   # ... and never present
    lab[0] = 0                                        # first sample: background only
    target = torch.nn.functional.one_hot(lab, C).movedim(-1, 1).float()
    lg = logits.to(dt)
    got = M.compute_dice(lg, target)
    ref = O.compute_dice(lg.float().cpu(), target.cpu())
    assert abs(got - ref) < 1e-6, (got, ref)
    conf = M.batch_confusion(lg, target)
    pred = lg.float().argmax(1)
    ref_conf = torch.zeros(B, C, C, dtype=torch.int64)
    for b in range(B):
        idx = (lab[b].reshape(-1) * C + pred[b].reshape(-1)).cpu()
        ref_conf[b] = torch.bincount(idx, minlength=C * C).view(C, C)
    assert torch.equal(conf.cpu(), ref_conf)
    rd = M.RunningDice(labels=list(range(C)), ignore_label=0)
    rd.update(lg, target)
    rd.update(lg[:1], target[:1])                     # all-background ground truth: skipped like the reference
    rd2 = M.RunningDice(labels=list(range(C)), ignore_label=0)
    rd2.update_matrix(lab.cpu().numpy(), pred.cpu().numpy())
    cm = ref_conf.sum(0).numpy()
    if not (lab != 0).any().item():
        cm = np.zeros_like(cm)                        # ground truth entirely the ignore label: the batch is skipped
    assert np.array_equal(rd.overall_confusion_matrix.cpu().numpy(), cm)
    assert np.array_equal(rd2.overall_confusion_matrix.cpu().numpy(), cm)
    mean, dl = rd.compute_dice()
    inter = np.diag(cm); union = cm.sum(0) + cm.sum(1)
    iou = (2 * inter + 1e-5) / (union.astype(np.float32) + 1e-5)
    assert abs(mean - float(np.mean(iou[1:]))) < 1e-6 and dl == [round(float(c), 4) for c in iou]


@pytest.mark.parametrize("C,ncls,V", [(32, 2, 4097), (64, 2, 729), (128, 4, 200), (16, 2, 130), (256, 3, 77)])
def test_instnorm_apply_fused_with_head_matches_separate_kernels(C, ncls, V):
    """hdf_instnorm_apply_head (one pass) against hdf_instnorm_apply followed by hdf_head_fwd: the activation is bit-identical,
    the logits differ only by the fp32 summation order of the 1x1x1 head (then bf16 rounding); ragged row counts exercise
    the masked tail where whole warps still take part in the shuffles."""
    torch.manual_seed(C + ncls)
    N = 2
    y = torch.randn(N, V, 1, 1, C, device=DEV).to(torch.bfloat16)
    assert ops.instnorm_apply_head_supported(y, ncls)
    mean, rstd = ops.instnorm_stats(y)
    gm, bt = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    hw, hb = torch.randn(ncls, C, 1, 1, 1, device=DEV) / C ** 0.5, torch.randn(ncls, device=DEV)
    buf = torch.zeros(N, V, 1, 1, C + 8, dtype=torch.bfloat16, device=DEV)
    out = buf[..., :C]
    logits = ops.instnorm_apply_head(y, mean, rstd, gm, bt, out, hw, hb, relu=True)
    ref_out = torch.empty_like(y)
    ops.instnorm_apply(y, mean, rstd, gm, bt, ref_out, relu=True)
    ref_logits = ops.head_fwd(ref_out, hw, hb)
    assert torch.equal(out, ref_out) and buf[..., C:].abs().max().item() == 0
    assert logits.shape == ref_logits.shape
    assert rel(logits.float(), ref_logits.float()) < 1e-2
    exact = torch.einsum("nvc,kc->nkv", ref_out.view(N, V, C).float(), hw.view(ncls, C)) + hb.view(1, ncls, 1)
    assert rel(logits.float().view(N, ncls, V), exact) < 1e-2
