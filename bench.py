#!/usr/bin/env python
"""Headline benchmark: HDenseFormer_32 3D training step at 2x144^3 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement of the reference
                                                             # (oracle/) timed on the host cores, rank 0 only

Prints ONE JSON line (rank 0).  A "step" = forward (bf16) + Dice/CE deep-supervision loss + backward + gradient
all-reduce (N>1) + Adam step on one synthetic batch of `--batch` volumes per GPU.
  value : volumes/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API (trainer.train_step) from PINNED HOST batches, H2D inside the timed
          region and a D2H read of the loss every step
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FWD_GF_PER_SAMPLE = {(144, 144, 144): 1410.6, (96, 96, 96): 417.6}   # SURVEY.md 8d (2*MAC, conv+linear+bmm)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def cpu_reference_step_time(size, batch, td, steps, warmup, budget_s):
    """Times the CPU restatement of the reference (oracle/hdf_oracle.py): fwd + DeepSuperloss(CEPlusDice) + backward,
    fp32, train-mode dropout on, all host threads.  Returns (seconds/step, steps actually timed, threads)."""
    from oracle import hdf_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    shapes = O.param_shapes(2, 2, 32, size, td)
    sd = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(shapes, seed=0).items()}
    x, t = O.synth_petct(batch, size, seed=0), O.synth_label(batch, 2, size, seed=0)
    g = torch.Generator().manual_seed(0)

    def one():
        for v in sd.values():
            v.grad = None
        outs = O.forward(sd, x, td, dropout_p=0.5, generator=g)
        O.deep_super_loss(outs, t, ignore_index=0).backward()

    t0 = time.perf_counter()
    one()
    first = time.perf_counter() - t0
    n_warm = min(warmup, 1 if first * (warmup + steps) > budget_s else warmup)
    for _ in range(max(0, n_warm - 1)):
        one()
    n = max(1, min(steps, int(budget_s / max(first, 1e-3)) - n_warm))
    t0 = time.perf_counter()
    for _ in range(n):
        one()
    return (time.perf_counter() - t0) / n, n, threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2, help="volumes per GPU per step (reference BATCH_SIZE=2, config.py:77)")
    ap.add_argument("--size", type=int, nargs=3, default=[144, 144, 144])
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--modalities", type=int, default=2, help="input channels (2 = PET/CT headline config; 3/4 = MR configs)")
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--fp32", action="store_true", help="exact fp32 path instead of bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam(fused=True) instead of the library's FusedAdam")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    a = ap.parse_args()
    size = tuple(a.size)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(a.warmup, int(os.environ.get("HDF_BENCH_MIN_WARMUP", 3)))   # timing rule: W >= 3 (env override only for ncu runs)
    kind = "PET/CT" if a.modalities == 2 else f"{a.modalities}-modality MR"
    workload = f"HDenseFormer_32 3D train step, {kind} {a.batch}x{a.modalities}x{size[0]}x{size[1]}x{size[2]} per GPU, " \
               f"td={a.depth}, " + (f"n_cls={a.classes}, " if a.classes != 2 else "") + "DeepSuperloss(CEPlusDice), Adam"
    gf_step = 3.0 * FWD_GF_PER_SAMPLE.get(size, 1410.6 * (size[0] * size[1] * size[2]) / 144 ** 3)   # per sample

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return
        sec, n, threads = cpu_reference_step_time(size, 1, a.depth, a.steps, a.warmup, budget_s=150.0)
        v = 1.0 / sec
        print(json.dumps({
            "impl": "reference", "metric": "3D train volumes/sec @2x144^3", "value": v, "unit": "volumes/s", "n_gpus": a.gpus,
            "steps": n, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "CPU restatement of the reference (oracle port), train mode"},
            "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": threads, "kind": "port",
                             "sample": f"{n} step(s) of 1x2x{size[0]}^3 fwd+loss+bwd fp32 on {threads} host threads"},
            "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    # stdout carries exactly ONE JSON line: anything native libraries print there meanwhile (NCCL's version banner)
    # is sent to stderr by pointing fd 1 at fd 2 until the result line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    from oracle import hdf_oracle as O   # only for synthetic data generation + the cpu_baseline leg
    from hdenseformer_b200 import _C, trainer as T
    from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
    from hdenseformer_b200.models import HDenseFormer_32

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/hdf_bench_nccl_%h_%p.log")   # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = HDenseFormer_32(a.modalities, a.classes, size, a.depth).to(dev).train()
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("position_embeddings"):
                p.normal_(0, 0.02)
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    decay = [p for n_, p in net.named_parameters() if p.dim() > 1 and not n_.endswith(".bias")]
    no_decay = [p for n_, p in net.named_parameters() if not (p.dim() > 1 and not n_.endswith(".bias"))]
    use_graph = not a.no_graph and (world == 1 or os.environ.get("HDF_BENCH_GRAPH_MULTI", "1") == "1")
    if a.torch_adam:
        opt = torch.optim.Adam([{"params": decay, "weight_decay": 1e-4}, {"params": no_decay, "weight_decay": 0.0}], lr=1e-3,
                               fused=True, capturable=use_graph)   # grouping of trainer.py:812-819
    else:
        from hdenseformer_b200.optim import FusedAdam              # one launch over the gradient arena, same grouping
        opt = FusedAdam(net, lr=1e-3, weight_decay=1e-4)
    dp = T.DataParallelTrainer(net, crit, opt, use_bf16=not a.fp32)
    nb = 2   # two distinct pinned batches, alternated
    synth_x = (lambda sd_: O.synth_petct(a.batch, size, seed=sd_)) if a.modalities == 2 else \
              (lambda sd_: O.synth_mr(a.batch, a.modalities, size, seed=sd_))
    host = [(synth_x(rank * 10 + i).pin_memory(), O.synth_label(a.batch, a.classes, size, seed=rank * 10 + i).pin_memory())
            for i in range(nb)]
    devb = [(x.to(dev), t.to(dev)) for x, t in host]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    graphed = None
    if use_graph:
        try:
            graphed = T.GraphedTrainStep(net, crit, opt, devb[0][0], devb[0][1], use_bf16=not a.fp32)
            ok = torch.tensor([1], device=dev)
        except Exception as e:   # keep the benchmark alive; the mode actually used is reported in config.launch
            print(f"[bench] rank {rank}: CUDA-graph capture failed, falling back to eager launches: {e!r}", file=sys.stderr)
            graphed = None
            ok = torch.tensor([0], device=dev)
        if world > 1:            # all ranks must agree on the launch mode (collectives inside / outside the graph)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                graphed = None
    stepper = graphed.step if graphed is not None else dp.step
    step_dev = lambda i: stepper(*devb[i % nb])

    # e2e: every step's batch comes from pinned host memory inside the timed region; with the graphed stepper the copy of
    # step i+1 is issued on a copy stream so that it overlaps step i's kernels.  Every step's loss is read back to the
    # host: the 4-byte D2H copy is enqueued right behind the step and consumed one step later (the usual way a training
    # loop logs its loss without draining the GPU queue every step); the last one is drained inside the timed region.
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = []
    losses_read = []

    def drain_one():
        buf, ev = pending.pop(0)
        ev.synchronize()
        losses_read.append(float(buf))     # host read of that step's result

    def step_e2e(i):
        loss = stepper(*host[i % nb])
        if graphed is not None:
            graphed.prefetch(*host[(i + 1) % nb])
        buf = loss_host[i % 2]
        buf.copy_(loss.detach().float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending.append((buf, ev))
        if len(pending) > 1:
            drain_one()

    def finish_e2e():
        while pending:
            drain_one()

    for i in range(warmup):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, a.steps)
    # kernels of libhdf_b200 per step: counted on one eager step (a graph replay submits the same kernels in one call)
    l0 = _C.load().hdf_launch_count()
    dp.step(*devb[0])
    torch.cuda.synchronize()
    launches = (_C.load().hdf_launch_count() - l0) * a.steps
    clocks = sampler.stop() if rank == 0 else None
    for i in range(2):
        step_e2e(i)
    finish_e2e()
    losses_read.clear()
    ms_e2e = timed(step_e2e, a.steps, finish_e2e)
    assert len(losses_read) == a.steps and all(v == v for v in losses_read), "e2e: every step's loss must reach the host"

    vols = a.batch * world * a.steps
    value = vols / (ms / 1e3)
    e2e = vols / (ms_e2e / 1e3)
    pk = peaks()

    # ---- dominant kernel: tcgen05 conv of block_1_1_right (64 -> 32 at full resolution), timed alone with CUDA events
    roof = None
    if not a.fp32:
        from hdenseformer_b200 import ops
        D, H, W = size
        xin = torch.randn(a.batch, D, H, W, 64, device=dev).to(torch.bfloat16)     # 2x382 MB at 144^3: exceeds L2
        wgt = torch.randn(32, 64, 3, 3, 3, device=dev) * 0.02
        yout = torch.empty(a.batch, D, H, W, 32, dtype=torch.bfloat16, device=dev)
        wp = ops.tc_pack(wgt, 64, 32, 27, 64 * 27, False)
        for _ in range(3):
            ops.tc_conv3d_fwd(xin, wp, None, yout)
        reps = 10
        kms = timed(lambda i: ops.tc_conv3d_fwd(xin, wp, None, yout), reps) / reps
        flops = 2.0 * a.batch * D * H * W * 27 * 64 * 32
        ach = flops / (kms / 1e3) / 1e12
        # DRAM bytes per launch of this kernel from `ncu --set full` (profiles/r1_ncu_conv_fwd_b11r_final.txt:
        # dram__bytes_read.sum 764.6 MB + dram__bytes_write.sum 365.8 MB; algorithmic 764.4 + 382.2 MB), same shape only
        traffic = 1.1304e9 if (a.batch == 2 and size == (144, 144, 144)) else None
        roof = {"bound": "tensor", "kernel": "tc_conv_fwd_kernel (block_1_1_right: 64->32 @ full res)", "achieved": ach,
                "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"], "traffic": traffic,
                "traffic_unit": "bytes/launch (ncu dram read+write)", "algorithmic_flop_per_launch": flops,
                "peak_source": pk["src"] + " burst (kernel timed alone)", "ms_per_launch": kms,
                "step_frac_of_sustained_peak": (gf_step * value / 1e3) / pk["tf_sustained"]}
        del xin, yout

    if rank == 0:
        out = {
            "metric": "3D train volumes/sec @2x144^3", "value": value, "unit": "volumes/s", "n_gpus": world, "steps": a.steps,
            "warmup": warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.fp32 else "bf16", "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": a.batch, "global_batch": a.batch * world,
                       "parallelism": f"dp{world}", "launch": "cuda_graph_replay" if graphed is not None else "eager", "l2": "inputs_exceed_l2 (activations >> 126 MB per step)",
                       "algorithmic_tflop_per_volume": gf_step / 1e3, "achieved_tflops": gf_step * value / 1e3},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "volumes/s", "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(sum(x.numel() * 4 + t.numel() * 4 for x, t in host[:1])), "d2h_bytes_per_step": 4,
                    "readback": "every step's loss is copied to pinned host memory and read one step later (last one inside the timed region)"},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if world == 1 and not a.no_cpu_baseline:
            sec, n, threads = cpu_reference_step_time((96, 96, 96), 1, a.depth, 3, 1, budget_s=25.0)
            out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "volumes/s", "cores": threads, "kind": "port",
                                   "sample": f"{n} step(s) of BASELINE config[0] (1x2x96^3 fwd+loss+bwd, fp32, train mode) on "
                                             f"{threads} host threads; oracle/hdf_oracle.py"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        # a captured graph holds NCCL kernels; tearing the communicator down underneath it can hang, so leave
        # together and skip interpreter teardown
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
