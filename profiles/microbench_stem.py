"""First convolution (2 -> 32 @ 2 x 144^3) alone: fused gather kernels vs the im2col + GEMM path, CUDA-event timed."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdenseformer_b200 import ops

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for cin, cout, size, N in [(2, 32, (144, 144, 144), 2), (4, 32, (128, 128, 128), 2), (3, 32, (32, 384, 384), 2)]:
    x = torch.randn(N, cin, *size, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / math.sqrt(27 * cin)
    y = torch.empty(N, *size, cout, dtype=torch.bfloat16, device="cuda")
    g = torch.randn(N, *size, cout, device="cuda").to(torch.bfloat16)
    dw = torch.empty_like(w)
    si = ops.StemInput(x)
    flop = 2.0 * N * size[0] * size[1] * size[2] * 27 * cin * cout
    tf = timeit(lambda: ops.stem_conv_fwd(si, w, y)); tw = timeit(lambda: ops.stem_conv_wgrad(si, g, dw))
    ti = timeit(lambda: ops.stem_im2col(x)); xcol = ops.stem_im2col(x)
    tfo = timeit(lambda: ops.stem_conv_fwd(xcol, w, y)); two = timeit(lambda: ops.stem_conv_wgrad(xcol, g, dw))
    byts = x.numel() * 4 + y.numel() * 2
    print(f"{cin}->{cout} @ {N}x{size}: fused fwd {tf:.3f} ms ({byts/tf/1e6:.0f} GB/s algorithmic, {flop/tf/1e9:.0f} TF/s)  fused wgrad {tw:.3f} ms | "
          f"im2col {ti:.3f} + gemm fwd {tfo:.3f} ms, wgrad {two:.3f} ms")
