import os, sys
sys.path.insert(0, "/root/repo")
import torch
from hdenseformer_b200 import ops
dev = "cuda"; ops.ensure_init(torch.zeros(1, device=dev))
B, N, H = 2, 729, 8
R = B * N
qkv = torch.randn(R, 96, device=dev)
for _ in range(3):
    o, lse = ops.attention_fwd(qkv, B, N, H, 0.5)
    do = torch.randn_like(o)
    ops.attention_bwd(qkv, o, do, lse, B, N, H, 0.5)
    A = torch.randn(R, 224, device=dev); W = torch.randn(32, 224, device=dev) * 0.1; bias = torch.zeros(32, device=dev)
    out = torch.empty(R, 32, device=dev)
    ops.gemm(A, W, True, out, bias=bias)
torch.cuda.synchronize()
