"""Generate the golden vectors that pin oracle/hdf_oracle.py to the unmodified reference.

Run in the build container only (needs /root/reference, which does not travel):

    python tests/golden/make_golden.py

It imports the reference's own models/HDenseFormer.py and loss/*.py, loads the
deterministic synthetic parameters from oracle.hdf_oracle.synth_state_dict into the
reference module, runs the reference in eval() (dropout off) with grad
enabled, and writes small .npz/.json fixtures next to this script.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from models.HDenseFormer import HDenseFormer  # noqa: E402  (reference)
from models.HDenseFormer_2D import HDenseFormer_2D  # noqa: E402  (reference)
from loss.combine_loss import CEPlusDice, DeepSuperloss  # noqa: E402  (reference)
from loss.dice_loss import DiceLoss  # noqa: E402
from loss.cross_entropy import CrossentropyLoss  # noqa: E402

from oracle import hdf_oracle as O  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)


def model_case(name, in_ch, n_cls, nf, size, td, batch):
    shapes = O.param_shapes(in_ch, n_cls, nf, size, td)
    cls = HDenseFormer if len(size) == 3 else HDenseFormer_2D
    ref = cls(in_ch, n_cls, nf, image_size=size, transformer_depth=td)
    ref_sd = ref.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()), "state_dict key order/name mismatch"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    sd = O.synth_state_dict(shapes, seed=7)
    ref.load_state_dict(sd)
    ref.eval()
    x = O.synth_petct(batch, size, seed=1) if in_ch == 2 else O.synth_mr(batch, in_ch, size, seed=1)
    tgt = O.synth_label(batch, n_cls, size, seed=1)
    outs = ref(x)
    crit = DeepSuperloss(CEPlusDice(weight=None, ignore_index=0))
    loss = crit(outs, tgt)
    loss.backward()
    grads = {k: p.grad for k, p in ref.named_parameters()}
    gstat = {k: [float(g.double().norm()), float(g.double().sum())] for k, g in grads.items()}
    keep = [k for k in grads if grads[k].numel() <= 4096 and ("blocks.0.0.layers.1" in k or "block_4" in k
                                                                 or "conv1x1" in k or "up3" in k
                                                                 or k.endswith("patch_embeddings.bias"))]
    arrs = {f"out{i}": o.detach().numpy() for i, o in enumerate(outs)}
    arrs["loss"] = np.array(loss.item(), dtype=np.float64)
    for k in keep:
        arrs["grad:" + k] = grads[k].numpy()
    np.savez(os.path.join(HERE, f"{name}.npz"), **arrs)
    meta = dict(in_channels=in_ch, n_cls=n_cls, n_filters=nf, image_size=list(size), transformer_depth=td,
                batch=batch, param_seed=7, data_seed=1, shapes={k: list(v) for k, v in shapes.items()},
                grad_stats=gstat, n_tensors=len(shapes), n_params=int(sum(np.prod(v) for v in shapes.values())))
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(meta, f)
    print(name, "loss", loss.item(), "tensors", len(shapes), "params", meta["n_params"])


def loss_case():
    g = torch.Generator().manual_seed(11)
    arrs = {}
    for tag, C in (("c3", 3), ("c2", 2), ("c4", 4)):
        p = torch.randn((2, C, 8, 8, 8), generator=g) * 2
        lab = torch.randint(0, C, (2, 8, 8, 8), generator=g)
        t = torch.stack([(lab == c) for c in range(C)], 1).float()
        w = torch.rand((C,), generator=g) + 0.5
        arrs[f"{tag}_p"], arrs[f"{tag}_t"], arrs[f"{tag}_w"] = p.numpy(), t.numpy(), w.numpy()
        for ii, ig in (("ig0", 0), ("ignone", None), ("ig1", 1)):
            for wt, wv in (("w", w), ("nw", None)):
                pp = p.clone().requires_grad_(True)
                l = CEPlusDice(weight=wv, ignore_index=ig)(pp, t)
                l.backward()
                arrs[f"{tag}_{ii}_{wt}_loss"] = np.array(l.item())
                arrs[f"{tag}_{ii}_{wt}_grad"] = pp.grad.numpy()
        arrs[f"{tag}_dice_ig0"] = np.array(DiceLoss(ignore_index=0)(p, t).item())
        arrs[f"{tag}_ce"] = np.array(CrossentropyLoss()(p, t).item())
    # deep supervision on a 4-level pyramid
    C = 3
    lab = torch.randint(0, C, (2, 16, 16, 16), generator=g)
    t = torch.stack([(lab == c) for c in range(C)], 1).float()
    outs = [torch.randn((2, C, 16 >> i, 16 >> i, 16 >> i), generator=g).requires_grad_(True) for i in range(4)]
    l = DeepSuperloss(CEPlusDice(ignore_index=0))(outs, t)
    l.backward()
    arrs["ds_t"] = t.numpy()
    for i, o in enumerate(outs):
        arrs[f"ds_p{i}"] = o.detach().numpy()
        arrs[f"ds_g{i}"] = o.grad.numpy()
    arrs["ds_loss"] = np.array(l.item())
    np.savez(os.path.join(HERE, "loss_cases.npz"), **arrs)
    print("loss cases written")


if __name__ == "__main__":
    model_case("model_nf16_32cube", 2, 3, 16, (32, 32, 32), 4, 1)
    model_case("model_nf8_aniso", 3, 2, 8, (16, 32, 48), 8, 2)
    model_case("model2d_nf16_64x48", 3, 3, 16, (64, 48), 4, 2)        # HDenseFormer_2D (models/HDenseFormer_2D.py)
    loss_case()
