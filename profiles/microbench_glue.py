"""Per-kernel timing of the bandwidth-bound glue at the headline shapes (2 x 144^3, nf=32, bf16) against the HBM
roofline: achieved GB/s = algorithmic bytes (unique input + output bytes) / time.  Also the stem (im2col + GEMM first
layer).  CUDA events, operands larger than L2.  Run on the GPU box:  python profiles/microbench_glue.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdenseformer_b200 import ops

dev = "cuda"; ops.ensure_init(torch.zeros(1, device=dev))
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peak = json.load(open(pk)).get("hbm_gbs", 6555.5) if os.path.exists(pk) else 6555.5
B, S, C = 2, 144, 32
bf = torch.bfloat16

ONCE = os.environ.get("HDF_GLUE_ONCE") is not None      # one launch per kernel (for ncu captures)

def timeit(fn, reps=10):
    if ONCE:
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def row(name, ms, nbytes):
    print(f"{name:34s} {ms:8.3f} ms  {nbytes/1e6:9.1f} MB  {nbytes/ms/1e6:8.0f} GB/s  {nbytes/ms/1e6/peak:6.1%} of {peak:.0f}")

y = torch.randn(B, S, S, S, C, device=dev).to(bf)
nb = y.numel() * 2
mean, rstd = ops.instnorm_stats(y)
row("instnorm_stats 32ch@144", timeit(lambda: ops.instnorm_stats(y)), nb)
gm = torch.rand(C, device=dev) + 0.5; bt = torch.randn(C, device=dev)
out = torch.empty_like(y)
row("instnorm_apply 32ch@144", timeit(lambda: ops.instnorm_apply(y, mean, rstd, gm, bt, out, relu=True)), 2 * nb)
cat = torch.empty(B, S, S, S, 2 * C, dtype=bf, device=dev)
res = torch.randn(B, S, S, S, C, device=dev).to(bf)
row("instnorm_apply +res -> concat slice", timeit(lambda: ops.instnorm_apply(y, mean, rstd, gm, bt, cat[..., C:], residual=res, relu=True)), 3 * nb)
dout = torch.randn(B, S, S, S, C, device=dev).to(bf)
dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
row("instnorm_bwd (reduce+apply) 32ch@144", timeit(lambda: ops.instnorm_bwd(dout, y, mean, rstd, gm, bt, dg, db, relu=True)), 5 * nb)
p = torch.empty(B, S // 2, S // 2, S // 2, C, dtype=bf, device=dev)
row("maxpool2_fwd 144->72", timeit(lambda: ops.maxpool2_fwd(y, p)), nb + nb // 8)
dx = torch.empty_like(y)
row("maxpool2_bwd 72->144", timeit(lambda: ops.maxpool2_bwd(y, p, dx, False)), 2 * nb + nb // 8)
xs = torch.randn(B, S // 2, S // 2, S // 2, C, device=dev).to(bf)
row("upsample2_fwd 72->144", timeit(lambda: ops.upsample2_fwd(xs, out)), nb + nb // 8)
dxs = torch.empty_like(xs)
row("upsample2_bwd 144->72", timeit(lambda: ops.upsample2_bwd(dout, dxs)), nb + nb // 8)
w = torch.randn(2, C, device=dev) * 0.1; b = torch.zeros(2, device=dev)
lo = ops.head_fwd(y, w, b)
row("head_fwd 32->2 @144", timeit(lambda: ops.head_fwd(y, w, b)), nb + lo.numel() * 2)
g = torch.randn_like(lo)
dw = torch.zeros_like(w); dbb = torch.zeros_like(b)
row("head_bwd (wgrad+dgrad) @144", timeit(lambda: ops.head_bwd(g, y, w, dx, dw, dbb, False)), 2 * nb + g.numel() * 2)
del cat, res, dout, dx, p, xs, dxs, out
# ---- stem
x = torch.randn(B, 2, S, S, S, device=dev)
xcol = ops.stem_im2col(x)
row("stem_im2col 2ch -> Kp=64", timeit(lambda: ops.stem_im2col(x)), x.numel() * 4 + xcol.numel() * 2)
wc = torch.randn(C, 2, 3, 3, 3, device=dev) * 0.1
row("stem_conv_fwd (GEMM K=64,N=32)", timeit(lambda: ops.stem_conv_fwd(xcol, wc, y)), xcol.numel() * 2 + nb)
dwc = torch.empty_like(wc)
row("stem_conv_wgrad", timeit(lambda: ops.stem_conv_wgrad(xcol, y, dwc)), xcol.numel() * 2 + nb)
