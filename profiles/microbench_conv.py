"""Per-layer timing of the tcgen05 convolution kernels at the headline shapes (2 x 144^3, nf=32): forward,
input-gradient and weight-gradient of every 3x3x3 conv / transposed conv of the model, CUDA events, inputs
larger than L2 for the big layers.  Run on the GPU box:  python profiles/microbench_conv.py [--batch 2]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdenseformer_b200 import ops

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=2); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default=""); a = ap.parse_args()
dev = "cuda"; ops.ensure_init(torch.zeros(1, device=dev))
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1590.0
L = [  # name, Cin, Cout, spatial, mode
    ("block_1_1_left(pad16)", 16, 32, 144, 0), ("block_1_2_left", 32, 32, 144, 0), ("block_1_1_right", 64, 32, 144, 0),
    ("block_1_2_right", 32, 32, 144, 0), ("block_2_1_left", 32, 64, 72, 0), ("block_2_2_left", 64, 64, 72, 0),
    ("block_2_1_right", 128, 64, 72, 0), ("up3", 64, 32, 72, 0), ("block_3_1_left", 64, 128, 36, 0),
    ("block_3_2_left", 128, 128, 36, 0), ("block_3_1_right", 256, 128, 36, 0), ("up2", 128, 64, 36, 0),
    ("block_4_1_left", 128, 256, 18, 0), ("block_4_2_left", 256, 256, 18, 0), ("up1", 256, 128, 18, 0),
    ("deep_conv", 256, 256, 9, 0), ("upconv_1(T)", 64, 32, 144, 1), ("upconv_2(T)", 128, 64, 72, 1), ("upconv_3(T)", 256, 128, 36, 1),
]

def timeit(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}; totf = 0.0
print(f"{'layer':24s} {'Cin':>4s} {'Cout':>4s} {'sp':>4s}  {'fwd ms':>8s} {'TF/s':>6s}  {'dgrad ms':>8s} {'TF/s':>6s}  {'wgrad ms':>8s} {'TF/s':>6s}")
for name, ci, co, sp, mode in L:
    if a.only and a.only not in name: continue
    B = a.batch
    isp = sp // 2 if mode == 1 else sp
    x = torch.randn(B, isp, isp, isp, ci, device=dev).to(torch.bfloat16)
    y = torch.empty(B, sp, sp, sp, co, dtype=torch.bfloat16, device=dev)
    g = torch.randn(B, sp, sp, sp, co, device=dev).to(torch.bfloat16)
    dx = torch.empty_like(x)
    taps = 27 if mode == 0 else 27 / 8
    flops = 2.0 * B * sp ** 3 * taps * ci * co
    if mode == 0:
        w = torch.randn(co, ci, 27, device=dev) * 0.05
        wp = ops.tc_pack(w, ci, co, 27, ci * 27, False); wpd = ops.tc_pack(w, co, ci, ci * 27, 27, True)
        tf = timeit(lambda: ops.tc_conv3d_fwd(x, wp, None, y, 0), a.reps)
        td = timeit(lambda: ops.tc_conv3d_fwd(g, wpd, None, dx, 0), a.reps)
        dw = torch.empty_like(w)
        tw = timeit(lambda: ops.tc_conv3d_wgrad(x, g, dw, 27, ci * 27, 0), a.reps) if ops.tc_wgrad_supported(0, ci, co) else float("nan")
    else:
        w = torch.randn(ci, co, 27, device=dev) * 0.05
        wp = ops.tc_pack(w, ci, co, co * 27, 27, False); wpd = ops.tc_pack(w, co, ci, 27, co * 27, False)
        if ops.tc_convt_supported(ci, co):
            wct = ops.tc_convt_pack(w.view(ci, co, 3, 3, 3))
            tf = timeit(lambda: ops.tc_convt_fwd(x, wct, None, y), a.reps)
        else:
            tf = timeit(lambda: ops.tc_conv3d_fwd(x, wp, None, y, 1), a.reps)
        td = timeit(lambda: ops.tc_conv3d_fwd(g, wpd, None, dx, 2), a.reps)
        dw = torch.empty_like(w)
        tw = timeit(lambda: ops.tc_conv3d_wgrad(x, g, dw, co * 27, 27, 1), a.reps) if ops.tc_wgrad_supported(1, ci, co) else float("nan")
    tot["fwd"] += tf; tot["dgrad"] += td; tot["wgrad"] += tw; totf += flops
    print(f"{name:24s} {ci:4d} {co:4d} {sp:4d}  {tf:8.3f} {flops/tf/1e9:6.0f}  {td:8.3f} {flops/td/1e9:6.0f}  {tw:8.3f} {flops/tw/1e9:6.0f}")
    del x, y, g, dx
print(f"TOTAL fwd {tot['fwd']:.2f} ms  dgrad {tot['dgrad']:.2f} ms  wgrad {tot['wgrad']:.2f} ms ; conv GF/pass {totf/1e9:.0f} ; "
      f"fwd {totf/tot['fwd']/1e9:.0f} TF/s = {totf/tot['fwd']/1e9/peak:.1%} of {peak} (measured burst)")
