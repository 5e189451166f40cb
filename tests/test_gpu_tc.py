"""tcgen05 / TMA implicit-GEMM convolution (csrc/tc_conv.cu) against torch's fp32 conv on bf16-rounded operands.
Products of bf16 numbers are exact in fp32, so only the accumulation order differs: tolerance 2e-3 of max|y|
before the bf16 output rounding (2^-8 relative) -> 1e-2 overall."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hdenseformer_b200 import ops
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"

CASES = [
    # cin, cout, (D,H,W), N
    (16, 16, (8, 8, 8), 1),          # SW32, full tiles
    (32, 32, (16, 8, 16), 2),        # SW64
    (64, 64, (8, 16, 16), 1),        # SW128
    (32, 64, (9, 9, 9), 2),          # partial tiles in every dim
    (64, 32, (18, 18, 18), 1),
    (128, 128, (6, 10, 12), 1),      # 2 K-chunks
    (256, 128, (9, 9, 9), 1),        # 4 K-chunks
    (128, 256, (4, 6, 10), 2),       # N = 256 (TMEM 512 cols)
    (48, 80, (5, 7, 11), 1),         # KC = 16 with 3 chunks, odd N
    (256, 256, (9, 9, 9), 2),        # block_4_2_left / deep_conv at nf=32 (the benchmarked model): 9^3 token grid, B=2
    (256, 256, (18, 18, 18), 1),     # block_4_2_left @144^3
    (512, 256, (9, 9, 9), 1),        # deep_conv with 4 modalities (BASELINE config 4)
    (32, 32, (24, 20, 36), 2),       # full-resolution 32->32 layers (fold + kh-fold plan), ragged W
    (64, 32, (16, 12, 40), 1),       # block_1_1_right shape class
]


@pytest.mark.parametrize("cin,cout,size,N", CASES)
def test_tc_conv_fwd_matches_torch(cin, cout, size, N):
    ops.ensure_init(torch.zeros(1, device=DEV))
    assert ops.tc_supported(0, cin, cout)
    torch.manual_seed(cin * 1000 + cout)
    x = torch.randn(N, *size, cin + 16, device=DEV).to(torch.bfloat16)
    xs = x[..., 8:8 + cin]                                   # channel slice input (ld != C)
    w = (torch.randn(cout, cin, 3, 3, 3, device=DEV) / math.sqrt(27 * cin))
    b = torch.randn(cout, device=DEV)
    wq = w.to(torch.bfloat16).float()
    ref = F.conv3d(xs.float().permute(0, 4, 1, 2, 3), wq, b, padding=1).permute(0, 2, 3, 4, 1)
    buf = torch.zeros(N, *size, cout + 8, dtype=torch.bfloat16, device=DEV)
    y = buf[..., 8:]                                         # channel slice output
    wp = ops.tc_pack(w, cin, cout, 27, cin * 27, False)
    assert torch.equal(wp[5].float(), wq[:, :, 0, 1, 2])
    ops.tc_conv3d_fwd(xs, wp, b, y)
    torch.cuda.synchronize()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
    assert buf[..., :8].abs().max().item() == 0              # nothing written outside the slice
    # dgrad form: flipped taps, swapped channels == gradient of conv wrt its input
    g = torch.randn(N, *size, cout, device=DEV).to(torch.bfloat16)
    xr = xs.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    F.conv3d(xr, wq, None, padding=1).backward(g.float().permute(0, 4, 1, 2, 3))
    if ops.tc_supported(0, cout, cin):
        wpd = ops.tc_pack(w, cout, cin, cin * 27, 27, True)
        dx = torch.empty(N, *size, cin, dtype=torch.bfloat16, device=DEV)
        ops.tc_conv3d_fwd(g, wpd, None, dx)
        refdx = xr.grad.permute(0, 2, 3, 4, 1)
        err = ((dx.float() - refdx).abs().max() / refdx.abs().max()).item()
        assert err < 1e-2, err


WS_CASES = [
    # cin, (D,H,W), N, bias
    (32, (6, 8, 30), 1, False),        # one W tile exactly (30 output columns), partial H tiles
    (32, (5, 13, 37), 2, True),        # ragged everything, 2 samples, bias (UpConv up3 has one)
    (64, (7, 9, 64), 1, False),        # 128-byte rows, 3 W tiles
    (64, (4, 20, 31), 2, True),
    (32, (40, 16, 16), 1, False),      # long D column: several ring wrap-arounds and D segments
    (32, (3, 3, 3), 1, True),          # volume smaller than one tile
]


@pytest.mark.parametrize("cin,size,N,use_bias", WS_CASES)
def test_tc_ws_conv_matches_torch_and_old_kernel(cin, size, N, use_bias):
    """Weight-stationary tcgen05 kernel (csrc/tc_conv_ws.cu: weights in TMEM, voxels on the MMA N side, plane ring) against
    torch's fp32 conv on the bf16-rounded operands, and bit-for-bit-close to the first-generation kernel; forward and
    input-gradient forms, channel-slice operands."""
    import os
    ops.ensure_init(torch.zeros(1, device=DEV))
    cout = 32
    assert ops.tc_ws_supported(0, cin, cout)
    torch.manual_seed(cin + size[2])
    xb = torch.randn(N, *size, cin + 16, device=DEV).to(torch.bfloat16)
    xs = xb[..., 8:8 + cin]
    w = torch.randn(cout, cin, 3, 3, 3, device=DEV) / math.sqrt(27 * cin)
    b = torch.randn(cout, device=DEV) if use_bias else None
    wq = w.to(torch.bfloat16).float()
    ref = F.conv3d(xs.float().permute(0, 4, 1, 2, 3), wq, b, padding=1).permute(0, 2, 3, 4, 1)
    buf = torch.zeros(N, *size, cout + 32, dtype=torch.bfloat16, device=DEV)
    y = buf[..., 32:]
    wp = ops.tc_pack(w, cin, cout, 27, cin * 27, False)
    ops.tc_ws_conv3d_fwd(xs, wp, b, y)
    torch.cuda.synchronize()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
    assert buf[..., :32].abs().max().item() == 0
    # same operands through the first-generation kernel (voxels on M): both round the same fp32 sums to bf16
    os.environ["HDF_TC_NO_WS"] = "1"
    try:
        y_old = torch.empty(N, *size, cout, dtype=torch.bfloat16, device=DEV)
        ops.tc_conv3d_fwd(xs, wp, b, y_old)
    finally:
        os.environ["HDF_TC_NO_WS"] = "0"
    torch.cuda.synchronize()
    d = (y.float() - y_old.float()).abs().max().item() / ref.abs().max().item()
    assert d < 8e-3, d          # at most one bf16 ulp apart (different fp32 summation order)
    # input-gradient form of a Cin'=cout... layer whose dgrad has 32 output channels: dy has `cin` channels, dx 32
    g = torch.randn(N, *size, cin, device=DEV).to(torch.bfloat16)
    w2 = torch.randn(cin, cout, 3, 3, 3, device=DEV) / math.sqrt(27 * cout)      # Conv3d(32 -> cin) weight
    w2q = w2.to(torch.bfloat16).float()
    xr = torch.zeros(N, cout, *size, device=DEV, requires_grad=True)
    F.conv3d(xr, w2q, None, padding=1).backward(g.float().permute(0, 4, 1, 2, 3))
    wpd = ops.tc_pack(w2, cin, cout, cout * 27, 27, True)       # packed[tap][n = ci(32)][k = co(cin)] = w2[co][ci][26 - tap]
    dx = torch.empty(N, *size, cout, dtype=torch.bfloat16, device=DEV)
    ops.tc_ws_conv3d_fwd(g, wpd, None, dx)
    refdx = xr.grad.permute(0, 2, 3, 4, 1)
    err = ((dx.float() - refdx).abs().max() / refdx.abs().max()).item()
    assert err < 1e-2, err


def test_tc_conv_matches_simt_path_large():
    """72^3 x 32->32: thousands of tiles through the persistent scheduler, vs our own SIMT kernel."""
    ops.ensure_init(torch.zeros(1, device=DEV))
    torch.manual_seed(1)
    x = torch.randn(1, 72, 72, 72, 32, device=DEV).to(torch.bfloat16)
    w = torch.randn(32, 32, 3, 3, 3, device=DEV) / math.sqrt(27 * 32)
    y1 = torch.empty(1, 72, 72, 72, 32, dtype=torch.bfloat16, device=DEV)
    y2 = torch.empty_like(y1)
    ops.tc_conv3d_fwd(x, ops.tc_pack(w, 32, 32, 27, 32 * 27, False), None, y1)
    ops.conv3d_fwd(x, ops.conv_pack(w.to(torch.bfloat16).float(), 32, 32, 27, 32 * 27, False), None, y2, 0)
    torch.cuda.synchronize()
    err = ((y1.float() - y2.float()).abs().max() / y2.float().abs().max()).item()
    assert err < 1e-2, err


WG_CASES = [
    (16, 16, (8, 8, 8), 1),       # SW32, 8 taps per MMA group
    (32, 32, (16, 16, 24), 2),    # 4 taps per group, several slabs
    (64, 32, (9, 9, 9), 2),       # partial tiles; 14 groups in one pass (448 TMEM columns)
    (32, 64, (8, 12, 20), 1),
    (64, 64, (10, 16, 16), 1),    # 2 passes
    (128, 128, (6, 10, 12), 1),   # 7 passes
    (256, 128, (9, 9, 9), 1),     # two 128-channel groups per tap
    (128, 256, (4, 6, 10), 2),    # KV = 64 chunks, N = 256
    (256, 256, (9, 9, 9), 2),     # block_4_2_left / deep_conv at nf=32
    (256, 256, (18, 18, 18), 1),
    (512, 256, (9, 9, 9), 1),     # deep_conv with 4 modalities
    # plane-ring kernel (csrc/tc_wgrad_ws.cu: one operand has 32 channels)
    (32, 32, (20, 18, 50), 2),    # ragged H / W tiles, 2 samples
    (32, 32, (40, 8, 16), 1),     # long D column: ring wrap-arounds, several D segments
    (128, 32, (6, 10, 24), 1),    # P = dy, Q = x in 4 channel passes
    (32, 256, (4, 9, 17), 1),     # P = x, Q = dy in 8 channel passes, W not a multiple of 8
    (64, 32, (12, 12, 40), 2),    # block_1_1_right / up3 shape class
    (32, 32, (3, 3, 3), 1),       # volume smaller than one tile
]


@pytest.mark.parametrize("cin,cout,size,N", WG_CASES)
def test_tc_conv_wgrad_matches_torch(cin, cout, size, N):
    ops.ensure_init(torch.zeros(1, device=DEV))
    assert ops.tc_wgrad_supported(0, cin, cout)
    torch.manual_seed(cin * 7 + cout)
    xb = torch.randn(N, *size, cin + 8, device=DEV).to(torch.bfloat16)
    x = xb[..., 8:]
    gb = torch.randn(N, *size, cout + 16, device=DEV).to(torch.bfloat16)
    g = gb[..., :cout]
    w = torch.zeros(cout, cin, 3, 3, 3, device=DEV, requires_grad=True)
    F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, None, padding=1).backward(g.float().permute(0, 4, 1, 2, 3))
    dw = torch.full((cout, cin, 3, 3, 3), 3.0, device=DEV)
    ops.tc_conv3d_wgrad(x, g, dw, 27, cin * 27, 0, accumulate=False)
    torch.cuda.synchronize()
    err = ((dw - w.grad).abs().max() / w.grad.abs().max()).item()
    assert err < 1e-3, err
    ops.tc_conv3d_wgrad(x, g, dw, 27, cin * 27, 0, accumulate=True)
    err = ((dw - 2 * w.grad).abs().max() / w.grad.abs().max()).item()
    assert err < 2e-3, err


CT_CASES = [(32, 16, (4, 6, 8), 2), (64, 32, (9, 9, 9), 1), (128, 64, (5, 6, 7), 1), (256, 128, (4, 4, 6), 1), (64, 32, (18, 18, 18), 1)]


@pytest.mark.parametrize("cin,cout,size,N", CT_CASES)
def test_tc_conv_transpose_fwd_dgrad_wgrad(cin, cout, size, N):
    """nn.ConvTranspose3d k3 s2 p1 op1 on tensor cores: 8 output-parity classes (fwd), stride-2 TMA gather
    (dgrad and wgrad)."""
    ops.ensure_init(torch.zeros(1, device=DEV))
    torch.manual_seed(cin + cout)
    x = torch.randn(N, *size, cin, device=DEV).to(torch.bfloat16)
    w = torch.randn(cin, cout, 3, 3, 3, device=DEV) / math.sqrt(8 * cin)
    b = torch.randn(cout, device=DEV)
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    ref = F.conv_transpose3d(xr, wq, b, stride=2, padding=1, output_padding=1)
    osz = tuple(2 * s for s in size)
    buf = torch.zeros(N, *osz, cout + 8, dtype=torch.bfloat16, device=DEV)
    y = buf[..., :cout]
    assert ops.tc_supported(1, cin, cout)
    ops.tc_conv3d_fwd(x, ops.tc_pack(w, cin, cout, cout * 27, 27, False), b, y, mode=1)
    torch.cuda.synchronize()
    refl = ref.detach().permute(0, 2, 3, 4, 1)
    err = ((y.float() - refl).abs().max() / refl.abs().max()).item()
    assert err < 1e-2, err
    assert buf[..., cout:].abs().max().item() == 0
    g = torch.randn(N, *osz, cout, device=DEV).to(torch.bfloat16)
    ref.backward(g.float().permute(0, 4, 1, 2, 3))
    dx = torch.empty(N, *size, cin, dtype=torch.bfloat16, device=DEV)
    assert ops.tc_supported(2, cout, cin)
    ops.tc_conv3d_fwd(g, ops.tc_pack(w, cout, cin, 27, cout * 27, False), None, dx, mode=2)
    refdx = xr.grad.permute(0, 2, 3, 4, 1)
    err = ((dx.float() - refdx).abs().max() / refdx.abs().max()).item()
    assert err < 1e-2, err
    if ops.tc_wgrad_supported(1, cin, cout):
        dw = torch.zeros_like(w)
        ops.tc_conv3d_wgrad(x, g, dw, cout * 27, 27, 1)
        err = ((dw - wq.grad).abs().max() / wq.grad.abs().max()).item()
        assert err < 1e-3, err


def test_tc_first_layer_zero_padded_channels():
    """2-channel input zero-padded to 16 channels so the first conv and its weight gradient run on tensor cores."""
    ops.ensure_init(torch.zeros(1, device=DEV))
    torch.manual_seed(3)
    x = torch.randn(2, 2, 16, 16, 16, device=DEV)
    w = torch.randn(32, 2, 3, 3, 3, device=DEV) / math.sqrt(54)
    xcl = ops.ncdhw_to_cl(x, torch.bfloat16, pad_to=16)
    assert xcl.shape[-1] == 16 and xcl[..., 2:].abs().max().item() == 0
    y = torch.empty(2, 16, 16, 16, 32, dtype=torch.bfloat16, device=DEV)
    ops.tc_conv3d_fwd(xcl, ops.tc_pack(w, 16, 32, 27, 2 * 27, False, cin_valid=2), None, y)
    xq = xcl[..., :2].float().permute(0, 4, 1, 2, 3)
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv3d(xq, wq, None, padding=1)
    refl = ref.detach().permute(0, 2, 3, 4, 1)
    assert ((y.float() - refl).abs().max() / refl.abs().max()).item() < 1e-2
    g = torch.randn(2, 16, 16, 16, 32, device=DEV).to(torch.bfloat16)
    ref.backward(g.float().permute(0, 4, 1, 2, 3))
    tmp = torch.empty(32, 16 * 27, device=DEV)
    ops.tc_conv3d_wgrad(xcl, g, tmp, 27, 16 * 27, 0)
    dw = tmp[:, :54].reshape(32, 2, 3, 3, 3)
    assert ((dw - wq.grad).abs().max() / wq.grad.abs().max()).item() < 1e-3
    assert tmp[:, 54:].abs().max().item() == 0


@pytest.mark.parametrize("cin,cout,size,N", [(2, 32, (16, 16, 16), 2), (1, 16, (8, 12, 20), 1), (3, 32, (16, 24, 40), 1),
                                             (4, 32, (12, 20, 18), 2), (2, 64, (9, 10, 11), 1)])
def test_stem_im2col_gemm_matches_torch(cin, cout, size, N):
    """First conv (1..4 input channels) through the im2col + tcgen05 GEMM stem path: gathered matrix is exact,
    forward and weight gradient against torch's fp32 conv on the bf16-rounded operands."""
    ops.ensure_init(torch.zeros(1, device=DEV))
    assert ops.stem_supported(cin, cout)
    torch.manual_seed(cin * 7 + cout)
    x = torch.randn(N, cin, *size, device=DEV)
    w = torch.randn(cout, cin, 3, 3, 3, device=DEV) / math.sqrt(27 * cin)
    xcol = ops.stem_im2col(x)
    Kp = ops.stem_kp(cin)
    assert Kp % 64 == 0 and tuple(xcol.shape) == (N, *size, Kp)
    # the gathered matrix equals torch's unfold of the zero-padded, bf16-rounded input, k = tap*Cin + ci
    xq = x.to(torch.bfloat16).float()
    xp = F.pad(xq, (1, 1, 1, 1, 1, 1))
    D, H, W = size
    cols = []
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                cols.append(xp[:, :, kd:kd + D, kh:kh + H, kw:kw + W])      # [N, cin, D, H, W]
    ref_col = torch.stack(cols, dim=1).permute(0, 3, 4, 5, 1, 2).reshape(N, D, H, W, 27 * cin)
    assert torch.equal(xcol[..., :27 * cin].float(), ref_col)
    assert xcol[..., 27 * cin:].abs().max().item() == 0
    y = torch.empty(N, *size, cout + 8, dtype=torch.bfloat16, device=DEV)[..., :cout]   # channel-slice output
    ops.stem_conv_fwd(xcol, w, y)
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv3d(xq, wq, None, padding=1)
    refl = ref.detach().permute(0, 2, 3, 4, 1)
    assert ((y.float() - refl).abs().max() / refl.abs().max()).item() < 1e-2
    g = torch.randn(N, *size, cout, device=DEV).to(torch.bfloat16)
    ref.backward(g.float().permute(0, 4, 1, 2, 3))
    dw = torch.full((cout, cin, 3, 3, 3), 7.0, device=DEV)
    ops.stem_conv_wgrad(xcol, g, dw)
    assert ((dw - wq.grad).abs().max() / wq.grad.abs().max()).item() < 1e-3


@pytest.mark.parametrize("cin,cout,size,N", [(2, 32, (16, 16, 16), 2), (1, 16, (8, 12, 20), 1), (3, 32, (16, 24, 40), 1),
                                             (4, 32, (12, 20, 18), 2), (2, 64, (9, 10, 11), 1), (2, 32, (48, 40, 56), 2),
                                             (4, 64, (5, 7, 9), 3)])
def test_stem_fused_gather_matches_torch_and_im2col_path(cin, cout, size, N):
    """First conv without the im2col matrix (csrc/stem_tc.cu: producer-gathered operand tiles, K-major in the forward,
    MN-major in the weight gradient): against torch's fp32 conv on bf16-rounded operands, and against the im2col + GEMM
    path, which multiplies exactly the same bf16 values (fp32 accumulation order differs).  Sizes whose voxel count is not
    a multiple of the 128-voxel tile and every border case of the 27 taps are included."""
    ops.ensure_init(torch.zeros(1, device=DEV))
    assert ops.stem_fused_supported(cin, cout)
    torch.manual_seed(cin * 11 + cout)
    x = torch.randn(N, cin, *size, device=DEV)
    w = torch.randn(cout, cin, 3, 3, 3, device=DEV) / math.sqrt(27 * cin)
    si = ops.StemInput(x)
    assert tuple(si.shape) == (N, *size, ops.stem_kp(cin)) and si.dtype == torch.bfloat16
    buf = torch.zeros(N, *size, cout + 8, dtype=torch.bfloat16, device=DEV)
    y = buf[..., :cout]
    ops.stem_conv_fwd(si, w, y)
    xq = x.to(torch.bfloat16).float()
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv3d(xq, wq, None, padding=1)
    refl = ref.detach().permute(0, 2, 3, 4, 1)
    assert ((y.float() - refl).abs().max() / refl.abs().max()).item() < 1e-2
    assert buf[..., cout:].abs().max().item() == 0
    y_old = torch.empty(N, *size, cout, dtype=torch.bfloat16, device=DEV)
    xcol = ops.stem_im2col(x)
    ops.stem_conv_fwd(xcol, w, y_old)
    assert ((y.float() - y_old.float()).abs().max() / refl.abs().max()).item() < 1e-2
    g = torch.randn(N, *size, cout, device=DEV).to(torch.bfloat16)
    ref.backward(g.float().permute(0, 4, 1, 2, 3))
    dw = torch.full((cout, cin, 3, 3, 3), 7.0, device=DEV)
    ops.stem_conv_wgrad(si, g, dw)
    assert ((dw - wq.grad).abs().max() / wq.grad.abs().max()).item() < 1e-3
    dw2 = dw.clone()
    ops.stem_conv_wgrad(si, g, dw2, accumulate=True)
    assert torch.allclose(dw2, 2 * dw, rtol=1e-6, atol=0)
    # channel-slice dY (row stride != Cout), as in the engine
    gb = torch.zeros(N, *size, cout + 16, dtype=torch.bfloat16, device=DEV)
    gb[..., :cout] = g
    dw3 = torch.empty_like(dw)
    ops.stem_conv_wgrad(si, gb[..., :cout], dw3)
    assert torch.equal(dw3, dw)


@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 32), (40, 24), (256, 512)])
def test_weight_pack_matches_permute(cin, cout):
    """hdf_tc_pack_weights for the two torch layouts (conv [Cout][Cin][27], conv-transpose [Cin][Cout][27]), plain and
    tap-flipped, against a permute."""
    g = torch.Generator(device="cuda").manual_seed(3)
    w = torch.randn(cout, cin, 27, device="cuda", generator=g)
    ref = w.permute(2, 0, 1).to(torch.bfloat16)                                  # [tap][co][ci]
    assert torch.equal(ops.tc_pack(w, cin, cout, 27, cin * 27, False).view(27, cout, cin), ref)
    assert torch.equal(ops.tc_pack(w, cin, cout, 27, cin * 27, True).view(27, cout, cin), ref.flip(0))
    wt = torch.randn(cin, cout, 27, device="cuda", generator=g)                  # [ci][co][tap]
    reft = wt.permute(2, 1, 0).to(torch.bfloat16)
    assert torch.equal(ops.tc_pack(wt, cin, cout, cout * 27, 27, False).view(27, cout, cin), reft)
    assert torch.equal(ops.tc_pack(wt, cin, cout, cout * 27, 27, True).view(27, cout, cin), reft.flip(0))


@pytest.mark.parametrize("size,N", [((4, 4, 8), 1), ((5, 6, 7), 2), ((9, 9, 9), 1), ((12, 8, 16), 2), ((18, 18, 18), 2)])
def test_shift_major_transposed_conv_matches_torch_and_tap_major_kernel(size, N):
    """csrc/tc_convt.cu (8 shifted boxes, 8 parity classes in TMEM) for ConvTranspose3d(64 -> 32, k3, s2, p1, op1) =
    upconv_1 (reference models/HDenseFormer.py:215) against torch fp32 on bf16-rounded operands and against the tap-major
    kernel (mode 1 of tc_conv3d_fwd); partial tiles in every dim, channel-sliced output like the decoder's cat buffer."""
    ops.ensure_init(torch.zeros(1, device=DEV))
    torch.manual_seed(11)
    cin, cout = 64, 32
    assert ops.tc_convt_supported(cin, cout)
    x = torch.randn(N, *size, cin, device=DEV).to(torch.bfloat16)
    w = torch.randn(cin, cout, 3, 3, 3, device=DEV) / math.sqrt(8 * cin)
    b = torch.randn(cout, device=DEV)
    ref = F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), b, stride=2, padding=1,
                             output_padding=1).permute(0, 2, 3, 4, 1)
    osz = tuple(2 * s for s in size)
    buf = torch.zeros(N, *osz, 2 * cout, dtype=torch.bfloat16, device=DEV)
    y = buf[..., :cout]
    ops.tc_convt_fwd(x, ops.tc_convt_pack(w), b, y)
    y_old = torch.empty(N, *osz, cout, dtype=torch.bfloat16, device=DEV)
    ops.tc_conv3d_fwd(x, ops.tc_pack(w, cin, cout, cout * 27, 27, False), b, y_old, mode=1)
    torch.cuda.synchronize()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
    assert buf[..., cout:].abs().max().item() == 0                  # the other half of the cat buffer is untouched
    # same products, fp32 accumulation in a different order -> the two kernels agree to bf16 rounding of the output
    assert ((y.float() - y_old.float()).abs().max() / ref.abs().max()).item() < 1e-2
    # the second launch re-uses the TMEM / mbarrier protocol state from scratch: identical bits
    y2 = torch.empty(N, *osz, cout, dtype=torch.bfloat16, device=DEV)
    ops.tc_convt_fwd(x, ops.tc_convt_pack(w), b, y2)
    assert torch.equal(y2, y.contiguous())
