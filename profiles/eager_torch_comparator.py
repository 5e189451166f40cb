"""Like-for-like GPU comparator: the same training step (forward + DeepSuperloss(CEPlusDice) + backward +
fused Adam, batch 2 x 144^3, td=12, nf=32) written as plain eager PyTorch ops (cuDNN convolutions, torch attention math)
on the same B200, fp32 without TF32 and bf16 autocast.  The op graph is the oracle's functional restatement of the
reference (oracle/hdf_oracle.py; /root/reference itself is not present on the GPU box) with dropout off.
Evidence script only: nothing here is on the product path or in bench.py."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hdf_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = "cuda"
size, B, td = (144, 144, 144), 2, 12
shapes = O.param_shapes(2, 2, 32, size, td)
sd = {k: v.to(dev).requires_grad_(True) for k, v in O.synth_state_dict(shapes, seed=0).items()}
opt = torch.optim.Adam(list(sd.values()), lr=1e-3, fused=True)
x, t = O.synth_petct(B, size, seed=0).to(dev), O.synth_label(B, 2, size, seed=0).to(dev)

def step(bf16):
    opt.zero_grad(set_to_none=True)
    if bf16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = O.forward(sd, x, td)
    else:
        outs = O.forward(sd, x, td)
    loss = O.deep_super_loss([o.float() for o in outs], t, ignore_index=0)
    loss.backward()
    opt.step()
    return loss

res = {}
for name, bf16 in (("bf16_autocast", True), ("fp32_no_tf32", False)):
    for _ in range(2): step(bf16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n): step(bf16)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = {"ms_per_step": ms, "volumes_per_s": B / (ms / 1e3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
print(json.dumps({"what": "eager PyTorch (cuDNN) comparator, same step, same B200", "batch": B, "size": size, **res}))
