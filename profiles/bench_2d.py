"""HDenseFormer_2D_32 (reference config.py:34-37,77: PI-CAI22, 384 x 384, batch 24) training step: ours (bf16, flat volumes on
the 3-D kernels) vs the unmodified reference 2-D module run eagerly by PyTorch/cuDNN under bf16 autocast on the same GPU, and
the reference on the host CPU (small sample).  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss  # noqa: E402
from hdenseformer_b200.models import HDenseFormer_2D_32  # noqa: E402
from oracle import hdf_oracle as O  # noqa: E402


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, M, C, size, td = int(os.environ.get("HDF_2D_BATCH", "24")), 3, 2, (384, 384), 12
    dev = "cuda"
    torch.manual_seed(0)
    x = O.synth_mr(B, M, size, seed=1).to(dev)
    t = O.synth_label(B, C, size, seed=1).to(dev)
    net = HDenseFormer_2D_32(M, C, size, td).to(dev).train()
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = net(x)
        l = crit(o, t)
        opt.zero_grad(set_to_none=True)
        l.backward()
        opt.step()

    for _ in range(3):
        step()
    ms = timed(step, 5)
    out = {"what": f"HDenseFormer_2D_32({M},{C},{size},td={td}) train step, batch {B}, bf16", "eager_launches": {"ms_per_step": ms,
           "slices_per_s": B / (ms / 1e3)}}
    try:
        from hdenseformer_b200 import trainer as T
        gopt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True, capturable=True)
        gs = T.GraphedTrainStep(net, crit, gopt, x, t, use_bf16=True)
        for _ in range(2):
            gs.step(x, t)
        gms = timed(lambda: gs.step(x, t), 10)
        out["cuda_graph_replay"] = {"ms_per_step": gms, "slices_per_s": B / (gms / 1e3), "loss": float(gs.loss.item())}
        del gs, gopt
    except Exception as e:
        out["cuda_graph_replay"] = {"unavailable": repr(e)[:300]}
    try:
        from oracle import stage_ref
        import importlib
        assert stage_ref.load() is not None, "oracle/_ref not staged (run __graft_entry__.build() where /root/reference exists)"
        mod = importlib.import_module("models.HDenseFormer_2D")
        loss_mod = importlib.import_module("loss.combine_loss")
        rnet = mod.HDenseFormer_2D_32(M, C, size, td).to(dev).train()
        rcrit = loss_mod.DeepSuperloss(loss_mod.CEPlusDice(weight=None, ignore_index=0))
        ropt = torch.optim.Adam(rnet.parameters(), lr=1e-3, fused=True)

        def rstep():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                o = rnet(x)
            l = rcrit(o, t)
            ropt.zero_grad(set_to_none=True)
            l.backward()
            ropt.step()

        for _ in range(2):
            rstep()
        rms = timed(rstep, 3)
        out["gpu_eager_reference"] = {"ms_per_step": rms, "slices_per_s": B / (rms / 1e3)}
        del rnet, ropt
        torch.cuda.empty_cache()
        cnet = mod.HDenseFormer_2D_32(M, C, size, td).train()
        xc, tc = x[:2].cpu(), t[:2].cpu()
        t0 = time.time()
        l = rcrit(cnet(xc), tc)
        l.backward()
        cs = time.time() - t0
        out["cpu_reference"] = {"s_per_step_batch2": cs, "slices_per_s": 2 / cs, "threads": torch.get_num_threads()}
    except Exception as e:
        out["reference_unavailable"] = repr(e)[:200]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
