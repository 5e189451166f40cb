"""BASELINE config 5: sliding-window inference of one 2x224^3 volume, patch 144^3, step 72 -> 27 patches
(trainer.py:488-593).  Reports ms/volume (CUDA events, input volume in pinned host memory, mask left on device)
for the bf16 path, and the Dice of the bf16 mask against the exact fp32-path mask of the same weights.
Multi-GPU: torchrun --nproc-per-node N profiles/bench_sliding_window.py  (patches sharded, one all-reduce)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import hdf_oracle as O
from hdenseformer_b200 import trainer as T
from hdenseformer_b200.models import HDenseFormer_32

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
size = (144, 144, 144)
shapes = O.param_shapes(2, 2, 32, size, 12)
net = HDenseFormer_32(2, 2, size, 12)
net.load_state_dict(O.synth_state_dict(shapes, seed=0))
net = net.to(dev).eval()
vol = O.synth_petct(1, (224, 224, 224), seed=5)[0].pin_memory()
reps = 3
for _ in range(2):
    mask16 = T.inference_slidingwindow(net, vol, 2, size, (72, 72, 72), use_bf16=True, use_graph=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    mask16 = T.inference_slidingwindow(net, vol, 2, size, (72, 72, 72), use_bf16=True, use_graph=True)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
mask32 = T.inference_slidingwindow(net, vol, 2, size, (72, 72, 72), use_bf16=False)
if rank == 0:
    dice = O.mask_dice(mask16.cpu(), mask32.cpu(), 2)
    agree = (mask16 == mask32).float().mean().item()
    print(json.dumps({"metric": "sliding-window ms/volume", "value": ms.item(), "unit": "ms", "n_gpus": world,
                      "config": {"workload": "2x224^3 volume, patch 144^3, step 72, 27 patches, HDenseFormer_32 td=12, bf16"},
                      "dice_bf16_vs_fp32_path": dice, "voxel_agreement": agree,
                      "foreground_fraction": (mask32 > 0).float().mean().item()}), flush=True)
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
