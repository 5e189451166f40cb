"""One eager (no CUDA graph) training step of the headline workload between cudaProfilerStart/Stop, after 2 warm-up
steps, for `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`:
the launch list of exactly one steady-state step.  Numbers printed under ncu are not bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hdf_oracle as O          # synthetic data only
from hdenseformer_b200 import trainer as T
from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
from hdenseformer_b200.models import HDenseFormer_32

size, B = (144, 144, 144), 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = HDenseFormer_32(2, 2, size, 12).to(dev).train()
crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
dp = T.DataParallelTrainer(net, crit, opt, use_bf16=True)
x, t = O.synth_petct(B, size, seed=0).to(dev), O.synth_label(B, 2, size, seed=0).to(dev)
for _ in range(2):
    dp.step(x, t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
dp.step(x, t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
