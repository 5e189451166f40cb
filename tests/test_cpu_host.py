"""CPU-only checks: the C-ABI library loads and exports every symbol include/hdf_b200.h declares, the drop-in
surface keeps the reference's state_dict, host-side sliding-window / bucketing logic, and the no-fallback rule."""
import ctypes
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hdf_oracle as O
from hdenseformer_b200 import _C
from hdenseformer_b200 import trainer as T
from hdenseformer_b200.engine import Config, GradArena, backward_param_order
from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
from hdenseformer_b200.models import HDenseFormer, HDenseFormer_16, HDenseFormer_32
from hdenseformer_b200.models.HDenseFormer import param_table


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    decl = _C.declared_symbols()
    assert len(decl) >= 45
    lib = ctypes.CDLL(_C.LIB_PATH)
    missing = [k for k in decl if not hasattr(lib, k)]
    assert not missing, missing
    lib2 = _C.load()
    assert lib2.hdf_version() >= 100
    assert isinstance(_C.last_error(), str)


def test_state_dict_matches_reference_table(golden_dir):
    for name in ("model_nf16_32cube", "model_nf8_aniso"):
        meta = json.load(open(os.path.join(golden_dir, name + ".json")))
        m = HDenseFormer(meta["in_channels"], meta["n_cls"], meta["n_filters"], tuple(meta["image_size"]),
                         meta["transformer_depth"])
        sd = m.state_dict()
        assert list(sd.keys()) == list(meta["shapes"].keys())
        assert all(list(v.shape) == meta["shapes"][k] for k, v in sd.items())
        assert len(list(m.buffers())) == 0
        m.load_state_dict(O.synth_state_dict(O.param_shapes(meta["in_channels"], meta["n_cls"], meta["n_filters"],
                                                            tuple(meta["image_size"]), meta["transformer_depth"])))
    # headline config: 406 tensors, 11.074 M parameters
    t = param_table(2, 2, 32, (144, 144, 144), 12)
    assert len(t) == 406 and sum(int(np.prod(s)) for s in t.values()) == 11073992
    t24 = param_table(2, 2, 32, (144, 144, 144), 24)       # shipped default depth (config.py:120): 11.561 M
    assert len(t24) == 742 and sum(int(np.prod(s)) for s in t24.values()) == 11561288
    assert HDenseFormer_16(2, 2, (32, 32, 32), 4).n_filters == 16
    assert HDenseFormer_32(2, 2, (32, 32, 32), 4).n_filters == 32


def test_constructor_rejects_illegal_sizes_and_cpu_inputs():
    with pytest.raises(ValueError):
        HDenseFormer(3, 2, 8, (24, 384, 384), 4)       # BASELINE config 3 as written is illegal in the reference too
    m = HDenseFormer(2, 2, 8, (32, 32, 32), 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 2, 32, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        CEPlusDice(ignore_index=0)(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 2, 4, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        DeepSuperloss(CEPlusDice())([torch.zeros(1, 2, 4, 4, 4)], torch.zeros(1, 2, 4, 4, 4))


def test_backward_order_covers_all_parameters_and_arena_is_aligned():
    m = HDenseFormer(3, 2, 8, (32, 32, 32), 8)
    keys = list(m.state_dict().keys())
    order = backward_param_order(Config(3, 2, 8, (32, 32, 32), 8), keys)
    assert sorted(order) == sorted(keys) and order[0].startswith("conv1x1.") and order[-1].startswith("attns.0.")
    arena = GradArena(dict(m.named_parameters()), order)
    assert all(o % 32 == 0 for o in arena.offsets.values())
    assert all(arena.views[k].shape == p.shape for k, p in m.named_parameters())


def test_cal_steps_and_patch_sharding_match_oracle():
    for vol, patch, step in (((224,) * 3, (144,) * 3, (72,) * 3), ((144, 160, 300), (144,) * 3, (72,) * 3),
                             ((48, 40, 32), (32,) * 3, (16,) * 3)):
        assert T.cal_steps(vol, patch, step) == O.cal_steps(vol, patch, step)
    steps = T.cal_steps((224,) * 3, (144,) * 3, (72,) * 3)
    patches = T.enumerate_patches(steps)
    assert len(patches) == 27 and patches[1] == (0, 0, 40)
    shards = [T.shard_patches(patches, r, 8) for r in range(8)]
    assert sorted(sum(shards, [])) == sorted(patches) and [len(s) for s in shards] == [4, 4, 4, 3, 3, 3, 3, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _bucket_worker(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    b = T.GradBucketer(flat, min_bucket_elems=200)
    for end in (100, 250, 300, 900):
        b.notify(end)
    b.finish()
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    expect = sum(gathered) / world
    ok = torch.allclose(flat, expect, atol=1e-6) and b.last_ranges == [(0, 250), (250, 900), (900, 1000)] and b.ranges == []
    # sliding-window merge: each rank accumulates its own patches, all-reduce(sum) == serial result
    steps = T.cal_steps((20, 20, 20), (16,) * 3, (8,) * 3)
    agg = torch.zeros(20, 20, 20)
    for (x, y, z) in T.shard_patches(T.enumerate_patches(steps), rank, world):
        agg[x:x + 16, y:y + 16, z:z + 16] += 1
    dist.all_reduce(agg)
    ref = torch.zeros(20, 20, 20)
    for (x, y, z) in T.enumerate_patches(steps):
        ref[x:x + 16, y:y + 16, z:z + 16] += 1
    ok = ok and torch.equal(agg, ref)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_bucketer_and_patch_merge_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in ps]
    assert res == [(0, True), (1, True)]


def test_fused_adam_and_stem_have_no_cpu_path_and_pure_helpers_answer_without_a_gpu():
    """Entry points that need no device answer on the CPU box (planning helpers); the optimizer refuses a CPU model."""
    from hdenseformer_b200.optim import FusedAdam
    lib = _C.load()
    assert lib.hdf_stem_kp(1) == 64 and lib.hdf_stem_kp(2) == 64 and lib.hdf_stem_kp(3) == 128 and lib.hdf_stem_kp(4) == 128
    assert lib.hdf_stem_kp(5) == 0 and lib.hdf_stem_supported(2, 32) == 1 and lib.hdf_stem_supported(2, 48) == 0
    assert lib.hdf_adam_chunk() > 0 and lib.hdf_adam_table_bytes(3) == 3 * lib.hdf_adam_table_bytes(1)
    buf = ctypes.create_string_buffer(lib.hdf_adam_table_bytes(2))
    assert lib.hdf_adam_table_set(buf, 1, 4096, 128, 10, 0.5) == 0
    assert lib.hdf_adam_table_set(buf, 0, 0, 0, 10, 0.0) != 0 and "hdf_adam_table_set" in _C.last_error()   # null param
    m = HDenseFormer(2, 2, 8, (32, 32, 32), 4)
    with pytest.raises(RuntimeError):
        FusedAdam(m, lr=1e-3)


def test_2d_model_state_dict_matches_the_reference_module_table():
    """HDenseFormer_2D: keys / shapes / parameter count of the reference's 2-D module (golden meta written by
    tests/golden/make_golden.py from models/HDenseFormer_2D.py); construction needs no GPU, forward refuses the CPU."""
    import json
    import os
    import pytest
    import torch
    from hdenseformer_b200.models import HDenseFormer_2D, HDenseFormer_2D_16
    meta = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "model2d_nf16_64x48.json")))
    m = HDenseFormer_2D(meta["in_channels"], meta["n_cls"], meta["n_filters"], tuple(meta["image_size"]), meta["transformer_depth"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(meta["shapes"].keys())
    assert all(list(v.shape) == meta["shapes"][k] for k, v in sd.items())
    assert sum(v.numel() for v in sd.values()) == meta["n_params"]
    assert HDenseFormer_2D_16(3, 2, (64, 64), 4).n_filters == 16
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 48))
    with pytest.raises(ValueError):
        HDenseFormer_2D(3, 2, 16, (60, 48), 4)
