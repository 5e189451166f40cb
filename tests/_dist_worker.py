"""torchrun worker for tests/test_gpu_multi.py: data-parallel gradient equality and sharded sliding window."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import hdf_oracle as O  # noqa: E402
from hdenseformer_b200 import trainer as T  # noqa: E402
from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss  # noqa: E402
from hdenseformer_b200.models import HDenseFormer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    size, td, nf = (64, 64, 64), 4, 8     # 64^3: well-conditioned gradients (SURVEY 8c pitfall 2), oracle still takes seconds
    shapes = O.param_shapes(2, 2, nf, size, td)
    sd = O.synth_state_dict(shapes, seed=7)
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    x, t = O.synth_petct(world, size, seed=21), O.synth_label(world, 2, size, seed=21)

    # reference: the ORACLE's gradient of the whole batch (what nn.DataParallel's gathered loss differentiates,
    # trainer.py:228-229,369-374), computed on the CPU by every rank
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.deep_super_loss(O.forward(sdg, x, td), t, ignore_index=0).backward()
    ref_g = {k: v.grad.to(dev) for k, v in sdg.items()}

    net = HDenseFormer(2, 2, nf, size, td)
    if rank == 0:
        net.load_state_dict(sd)          # other ranks start from their own random init: the trainer must broadcast
    net = net.to(dev).eval()
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    dp = T.DataParallelTrainer(net, crit, opt, use_bf16=False, min_bucket_elems=1 << 12)
    loss = dp.step(x[rank:rank + 1], t[rank:rank + 1])
    torch.cuda.synchronize()
    worst, min_cos = 0.0, 1.0
    for k, p in net.named_parameters():
        den = ref_g[k].abs().max().item()
        if den < 1e-6 or k.endswith("double_conv.0.bias"):      # identically-zero true gradient (SURVEY 8c pitfall 1)
            continue
        worst = max(worst, (p.grad - ref_g[k]).abs().max().item() / den)
        a, b = p.grad.double().flatten(), ref_g[k].double().flatten()
        min_cos = min(min_cos, (a @ b).item() / max(a.norm().item() * b.norm().item(), 1e-300))
    assert worst < 1e-2 and min_cos > 0.9999, \
        f"rank {rank}: data-parallel gradient differs from the oracle's batch gradient: max rel {worst}, min cos {min_cos}"
    ranges = dp._bucketer.last_ranges
    assert len(ranges) >= 3, f"expected several gradient buckets in flight during backward, got {ranges}"
    assert ranges[0][0] == 0 and ranges[-1][1] == net._grad_arena().total
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])), ranges

    # sharded sliding window == serial sliding window
    vol = O.synth_petct(1, (96, 80, 64), seed=11)[0]
    mask, prob = T.inference_slidingwindow(net, vol, 2, size, (32, 32, 32), use_bf16=False, return_prob=True, patch_batch=2)
    dist.barrier()
    if rank == 0:
        ref_mask, ref_prob = O.sliding_window(lambda d: O.forward(sd, d, td)[0], vol, 2, size, (32, 32, 32))
        err = ((prob.cpu() - ref_prob[0]).abs().max() / ref_prob.abs().max()).item()
        assert err < 1e-4, err
        assert O.mask_dice(mask.cpu(), ref_mask, 2) > 0.9999
        print(f"DIST_OK world={world} grad_err={worst:.2e} min_cos={min_cos:.6f} buckets={len(ranges)} sw_err={err:.2e} loss={loss.item():.5f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
