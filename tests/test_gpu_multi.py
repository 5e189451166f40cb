"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped on the 1-GPU box): one process per GPU under torchrun,
NCCL gradient all-reduce by buckets == single-GPU batch gradient; sliding-window patches sharded across ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_and_sliding_window():
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dist_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", worker], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-8000:]
    try:      # keep the pass line next to the other run artefacts (gpurun_out/ travels back from the GPU box)
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r2_multi_gpu_pass.log"), "a") as f:
            f.write("".join(l + "\n" for l in r.stdout.splitlines() if "DIST_OK" in l))
    except OSError:
        pass
