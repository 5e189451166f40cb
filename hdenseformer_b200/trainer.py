"""Host-side drivers of the hot path, mirroring the reference's call sites.

  train_step                 trainer.py:361-380   (H2D, autocast forward, loss outside autocast, zero_grad,
                                                    backward, optimizer step; no per-step host sync)
  GradBucketer / DataParallelTrainer
                             trainer.py:228-229   (nn.DataParallel -> one process per GPU, NCCL all-reduce of the
                                                    flat gradient arena, bucket by bucket while backward runs)
  cal_steps                  trainer.py:595-618
  inference_slidingwindow    trainer.py:488-593   (patches sharded across ranks; analytic count map)
  compute_dice               trainer.py:919-945
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

# ----------------------------------------------------------------------------- sliding window (host logic)


def cal_steps(image_size: Sequence[int], patch_size: Sequence[int], step_size: Sequence[int]) -> List[List[int]]:
    """Window start offsets per axis (trainer.py:595-618)."""
    steps = []
    for dim in range(len(image_size)):
        if image_size[dim] <= patch_size[dim]:
            steps.append([0])
            continue
        max_step_value = image_size[dim] - patch_size[dim]
        num_steps = int(np.ceil(max_step_value / step_size[dim])) + 1
        actual = max_step_value / (num_steps - 1)
        steps.append([int(np.round(actual * i)) for i in range(num_steps)])
    return steps


def enumerate_patches(steps: List[List[int]]) -> List[Tuple[int, int, int]]:
    """Patch origins in the reference's loop order (x outermost, z innermost; trainer.py:530-541)."""
    return [(x, y, z) for x in steps[0] for y in steps[1] for z in steps[2]]


def shard_patches(patches: List[Tuple[int, int, int]], rank: int, world: int) -> List[Tuple[int, int, int]]:
    """Round-robin patch ownership: rank r takes patches r, r+world, ...."""
    return patches[rank::world]


class _GraphedForward:
    """Eval forward of one fixed-shape patch captured in a CUDA graph (static input / output buffers)."""

    def __init__(self, net, shape, use_bf16):
        dev = next(net.parameters()).device
        self.net, self.use_bf16 = net, use_bf16
        self.x = torch.zeros(shape, dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        if self.use_bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.net(self.x)[0]
        return self.net(self.x)[0]

    def __call__(self, data):
        self.x.copy_(data, non_blocking=True)
        self.graph.replay()
        return self.out


_FWD_CACHE = {}


def _graphed_forward(net, shape, use_bf16):
    # the graph bakes in parameter addresses, not values: it stays valid across optimizer steps / load_state_dict
    key = (id(net), tuple(shape), bool(use_bf16), tuple(p.data_ptr() for p in list(net.parameters())[:4]))
    if key not in _FWD_CACHE:
        while len(_FWD_CACHE) >= 4:          # bounded: graphs pin their static buffers and every saved activation
            _FWD_CACHE.pop(next(iter(_FWD_CACHE)))
        _FWD_CACHE[key] = _GraphedForward(net, shape, use_bf16)
    return _FWD_CACHE[key]


@torch.no_grad()
def inference_slidingwindow(net, image, n_cls: int, patch_size: Sequence[int], step_size: Sequence[int],
                            use_bf16: bool = False, group=None, return_prob: bool = False, use_graph: bool = False,
                            patch_batch: int = 1):
    """Sliding-window inference of one volume `image` [M, X, Y, Z] (numpy or tensor, host or device).

    Every rank of `group` (or the single process) evaluates its share of the patches and accumulates
    softmax probabilities into a private fp32 buffer; the buffers are summed with one all-reduce, then
    normalised by the analytic per-voxel window count and arg-maxed on device.  Returns the int64 mask
    [X, Y, Z] on the device (and the averaged probabilities if return_prob).

    `patch_batch` > 1 evaluates that many patches per forward call (the reference loops with batch 1,
    trainer.py:530-577; InstanceNorm, LayerNorm and attention are per-sample, so the results do not change, while a
    batch-1 forward is latency-bound on a B200)."""
    from . import ops
    dev = next(net.parameters()).device
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(image)
    image = image.float()
    M, X, Y, Z = image.shape
    for s, p in zip((X, Y, Z), patch_size):
        if s < p:
            raise ValueError(f"volume {(X, Y, Z)} smaller than patch {tuple(patch_size)}: position embeddings have a fixed "
                             "token count (models/HDenseFormer.py:119)")
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    steps = cal_steps((X, Y, Z), patch_size, step_size)
    mine = shard_patches(enumerate_patches(steps), rank, world)
    was_training = net.training
    net.eval()
    agg = torch.zeros((n_cls, X, Y, Z), dtype=torch.float32, device=dev)
    px, py, pz = patch_size
    # The reference slices every patch on the host and uploads it (trainer.py:543-549): 27 strided host copies + 27 H2D
    # transfers per volume.  Here the volume goes to the device once (90 MB for 2 x 224^3; from pinned memory the copy is
    # asynchronous) and the patches are strided device-side slices.
    if image.device != dev:
        image = image.to(dev, non_blocking=True)
    patch_batch = max(1, int(patch_batch))
    for g0 in range(0, len(mine), patch_batch):
        grp = mine[g0:g0 + patch_batch]
        data = torch.stack([image[:, x:x + px, y:y + py, z:z + pz] for (x, y, z) in grp])
        if use_graph:
            logits = _graphed_forward(net, (len(grp), M, px, py, pz), use_bf16)(data)
        elif use_bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = net(data)[0]
        else:
            logits = net(data)[0]
        for j, (x, y, z) in enumerate(grp):
            ops.sw_accumulate(logits[j:j + 1].contiguous(), agg, x, y, z)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM, group=group)
    mask = ops.sw_finalize(agg, steps, list(patch_size), normalise=return_prob)
    net.train(was_training)
    return (mask, agg) if return_prob else mask


def compute_dice(predict: torch.Tensor, target: torch.Tensor, ignore_index: int = 0, smooth: float = 1e-5) -> float:
    """Hard-mask Dice over foreground classes (trainer.py:891-945), computed on the device from one confusion-count kernel
    (hdenseformer_b200.metrics); the only host synchronisation is the final read of the scalar."""
    from . import metrics
    return metrics.compute_dice(predict, target, ignore_index, smooth)


# ----------------------------------------------------------------------------- training


def train_step(net, criterion, optimizer, data: torch.Tensor, target: torch.Tensor, use_bf16: bool = True):
    """One optimizer step, same order as trainer.py:366-380.  `data` / `target` may be pinned host tensors.
    Returns the loss tensor (device; reading it is the caller's sync)."""
    dev = next(net.parameters()).device
    data = data.to(dev, non_blocking=True)
    target = target.to(dev, non_blocking=True)
    if use_bf16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            output = net(data)
    else:
        output = net(data)
    loss = criterion(output, target)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss


class GraphedTrainStep:
    """One optimizer step (forward, loss, backward, optimizer) captured in a single CUDA graph and replayed per step:
    ~1800 kernel launches per step are submitted with one call, so the host never paces the GPU.  The batch shape is
    fixed; inputs are copied into static device buffers (H2D from pinned host memory is part of `step`).  Dropout
    masks still change every step (the seed counter is device-resident and advanced inside the graph).  The
    optimizer must be capturable (torch.optim.Adam(..., fused=True, capturable=True))."""

    def __init__(self, net, criterion, optimizer, example_data: torch.Tensor, example_target: torch.Tensor,
                 use_bf16: bool = True, warmup: int = 3):
        self.net, self.criterion, self.optimizer, self.use_bf16 = net, criterion, optimizer, use_bf16
        self._copy_stream, self._staged = None, None
        dev = next(net.parameters()).device
        self.x = torch.empty(example_data.shape, dtype=torch.float32, device=dev)
        self.t = torch.empty(example_target.shape, dtype=torch.float32, device=dev)
        self.x.copy_(example_data)
        self.t.copy_(example_target)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # capture the main path on a high-priority stream: the kernel nodes inherit it, so the block scheduler
        # places the big persistent convolution CTAs ahead of the transformer branch's small kernels (which are
        # captured from default-priority side streams) whenever both are pending
        hp = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get("HDF_NO_PRIORITY") is None else None
        with torch.cuda.graph(self.graph, stream=hp):
            self.loss = self._eager()

    def _eager(self):
        if self.use_bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = self.net(self.x)
        else:
            out = self.net(self.x)
        loss = self.criterion(out, self.t)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.optimizer.step()
        return loss

    def step(self, data: torch.Tensor, target: torch.Tensor):
        if self._staged is not None and self._staged[0] is data and self._staged[1] is target:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged[2])     # H2D of this batch was prefetched
            self.x.copy_(self._sx)
            self.t.copy_(self._st)
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(cur)      # the staging buffers may be overwritten from here on (NOT after the whole step)
            self._staged = None
        else:
            self.x.copy_(data, non_blocking=True)
            self.t.copy_(target, non_blocking=True)
        sync_lr = getattr(self.optimizer, "sync_lr", None)
        if sync_lr is not None:
            sync_lr()        # an LR scheduler writes param_groups[..]['lr'] on the host; the replayed step reads the device copy
        self.graph.replay()
        return self.loss

    def prefetch(self, data: torch.Tensor, target: torch.Tensor):
        """Start the host->device copy of the NEXT step's (pinned) batch on a copy stream so that it overlaps the
        current step's kernels; `step(data, target)` with the same tensors then only does a device-side copy."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.x.device)
            self._sx, self._st = torch.empty_like(self.x), torch.empty_like(self.t)
        cs = self._copy_stream
        # staging buffers are free once the previous step's device-side copy consumed them: wait for THAT copy, not for
        # the whole step that was enqueued behind it (otherwise the H2D transfer never overlaps the step's kernels)
        free_ev = getattr(self, "_staging_free", None)
        if free_ev is not None:
            cs.wait_event(free_ev)
        else:
            cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            self._sx.copy_(data, non_blocking=True)
            self._st.copy_(target, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._staged = (data, target, ev)


class GradBucketer:
    """All-reduces (average) contiguous ranges of a flat gradient buffer as soon as backward reports them final.

    `notify(end_offset)` launches an asynchronous all-reduce of flat[prev_end:end_offset]; `finish()` makes the
    current stream wait for all of them (no host block with NCCL).  Works with any backend (gloo in CPU tests)."""

    def __init__(self, flat: torch.Tensor, group=None, min_bucket_elems: int = 1 << 20):
        self.flat, self.group, self.min_bucket = flat, group, min_bucket_elems
        self.world = dist.get_world_size(group)
        self.prev = 0
        self.works = []
        self.ranges: List[Tuple[int, int]] = []
        self.last_ranges: List[Tuple[int, int]] = []

    def notify(self, end: int, force: bool = False):
        if end - self.prev < self.min_bucket and not force:
            return
        if end <= self.prev:
            return
        seg = self.flat[self.prev:end]
        if self.world > 1:
            if dist.get_backend(self.group) == "nccl":
                self.works.append(dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                w = dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self.works.append((w, seg))
        self.ranges.append((self.prev, end))
        self.prev = end

    def finish(self):
        self.notify(self.flat.numel(), force=True)
        for w in self.works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        self.works = []
        self.prev = 0
        self.last_ranges, self.ranges = self.ranges, []      # buckets issued during the step that just finished


class DataParallelTrainer:
    """One process per GPU; every rank holds a full replica (11 M parameters) and a batch shard.  Replaces
    nn.DataParallel (trainer.py:228-229): no per-step parameter broadcast, no output gather, gradients are
    averaged with NCCL over NVLink while the remaining backward kernels run."""

    def __init__(self, net, criterion, optimizer, use_bf16: bool = True, group=None, min_bucket_elems: int = 1 << 20):
        self.net, self.criterion, self.optimizer, self.use_bf16, self.group = net, criterion, optimizer, use_bf16, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.min_bucket = min_bucket_elems
        self._bucketer: Optional[GradBucketer] = None
        if self.world > 1:
            for p in net.parameters():          # replicas start identical (rank 0's weights)
                dist.broadcast(p.data, src=0, group=group)
            net.grad_sync = self._on_ready

    def _on_ready(self, key: str):
        arena = self.net._grad_arena()
        if self._bucketer is None or self._bucketer.flat is not arena.flat:
            self._bucketer = GradBucketer(arena.flat, self.group, self.min_bucket)
        idx = arena.order.index(key)
        end = arena.offsets[arena.order[idx + 1]] if idx + 1 < len(arena.order) else arena.total
        self._bucketer.notify(end)
        if idx + 1 == len(arena.order):
            self._bucketer.finish()

    def step(self, data: torch.Tensor, target: torch.Tensor):
        return train_step(self.net, self.criterion, self.optimizer, data, target, self.use_bf16)
