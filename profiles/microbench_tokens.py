"""Stand-alone (warm, nothing else on the GPU) durations of the token-branch kernels at the headline shape:
R = 2 x 729 tokens per modality, 32-wide layers, 8 heads of dim 4.  Kernel durations from CUPTI records."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdenseformer_b200 import ops

dev = "cuda"; ops.ensure_init(torch.zeros(1, device=dev))
B, N, H = 2, 729, 8
R = B * N

from torch.profiler import profile, ProfilerActivity

def timeit(fn, reps=20):
    """GPU-side duration (CUPTI kernel records, summed over the kernels of one call): the python wrappers allocate
    outputs and go through ctypes, so wall-clock / event timing of these ~10 us kernels would measure the host."""
    for _ in range(5): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps): fn()
        torch.cuda.synchronize()
    tot = sum(e.time_range.end - e.time_range.start for e in prof.events()
              if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None)
    return tot / reps

qkv = torch.randn(R, 96, device=dev)
o, lse = ops.attention_fwd(qkv, B, N, H, 0.5)
print(f"attention_fwd            {timeit(lambda: ops.attention_fwd(qkv, B, N, H, 0.5)):8.1f} us")
do = torch.randn_like(o)
print(f"attention_bwd (dq+dkv)   {timeit(lambda: ops.attention_bwd(qkv, o, do, lse, B, N, H, 0.5)):8.1f} us")
for K, Nn in ((128, 32), (224, 32), (32, 96), (256, 64), (64, 128)):
    A = torch.randn(R, K, device=dev); W = torch.randn(Nn, K, device=dev) * 0.1; bias = torch.zeros(Nn, device=dev)
    out = torch.empty(R, Nn, device=dev)
    print(f"gemm  [{R}x{K}] x [{Nn}x{K}]^T   {timeit(lambda: ops.gemm(A, W, True, out, bias=bias)):8.1f} us")
x = torch.randn(R, 32, device=dev); g = torch.ones(32, device=dev); b = torch.zeros(32, device=dev)
print(f"layernorm_fwd            {timeit(lambda: ops.layernorm_fwd(x, g, b)):8.1f} us")
P = {"q.1.fn.to_out.0.weight": torch.randn(32, 32, device=dev) * 0.1, "q.1.fn.to_out.0.bias": torch.zeros(32, device=dev),
     "q.2.norm.weight": g, "q.2.norm.bias": b, "q.2.fn.net.0.weight": torch.randn(64, 32, device=dev) * 0.1,
     "q.2.fn.net.0.bias": torch.zeros(64, device=dev), "q.2.fn.net.3.weight": torch.randn(32, 64, device=dev) * 0.1,
     "q.2.fn.net.3.bias": torch.zeros(32, device=dev)}
G = {k: torch.zeros_like(v) for k, v in P.items()}
fout = torch.empty(R, 32, device=dev)
ids = (1, 2, 3, 4, 5)
sv = ops.dct_c_fwd(o, x, P, "q.", fout, 0.5, 0, ids)
print(f"dct_c_fwd                {timeit(lambda: ops.dct_c_fwd(o, x, P, 'q.', fout, 0.5, 0, ids)):8.1f} us   (includes 12 torch.empty)")
dg2 = torch.randn(R, 32, device=dev)
print(f"dct_c_bwd (+reduce)      {timeit(lambda: ops.dct_c_bwd(dg2, o, sv, P, G, 'q.', 0.5, 0, ids)):8.1f} us")
