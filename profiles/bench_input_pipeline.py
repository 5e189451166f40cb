"""GPU input pipeline (csrc/prep.cu) vs the numpy/scipy restatement of the reference chain, one 2 x 176^3 PET/CT volume ->
2 x 144^3 patch (crop, PET/CT normalise, 'tr' warp, flip, one-hot).  Prints one JSON line."""
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdenseformer_b200 import data_utils as DU  # noqa: E402
from oracle import prep_oracle as PO  # noqa: E402


def main():
    torch.manual_seed(0)
    vshape, patch, ncls, M = (176, 176, 176), (144, 144, 144), 2, 2
    vol = torch.randn(M, *vshape) * 500
    vol[1] = torch.exp(torch.randn(*vshape))
    lab = (torch.rand(*vshape) > 0.97).float()
    res = DU.ResidentVolumes([{"image": vol, "label": lab}] * 2)
    chain = DU.Compose([DU.RandomCrop3D(patch), DU.PETandCTNormalize(), DU.RandomTranslationRotationZoom3D("tr", ncls),
                        DU.RandomFlip3D("hv"), DU.To_Tensor(ncls, M)])
    ds = DU.DataGenerator(res, num_class=ncls, transform=chain)
    bi = torch.empty(2, M, *patch, device="cuda"); bl = torch.empty(2, ncls, *patch, device="cuda")
    random.seed(0); np.random.seed(0)
    for _ in range(3):
        DU.collate_batch(ds, [0, 1], bi, bl)
    torch.cuda.synchronize()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(K):
        DU.collate_batch(ds, [0, 1], bi, bl)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.time() - t0) / K * 1e3
    ms = e0.elapsed_time(e1) / K
    V = patch[0] * patch[1] * patch[2]
    alg = 2 * V * 4 * ((M + 1) + (M + ncls))             # bytes per batch of 2: raw image + label read, outputs written
    # CPU: the restated reference chain on one sample (the reference runs it in DataLoader workers)
    img, lb = vol.numpy(), lab.numpy()
    random.seed(0); np.random.seed(0)
    origin, wm, fl = PO.draw_crop(img.shape, patch), PO.draw_trz("tr"), PO.draw_flip("hv")
    t0 = time.time()
    PO.pipeline(img, lb, patch, ncls, M, norm="petct", origin=origin, warp_mat=wm, flip_axis=fl)
    cpu_s = time.time() - t0
    print(json.dumps({"what": "input pipeline, batch of 2 x (2 x 176^3 -> 2 x 144^3), crop+PET/CT norm+'tr' warp+flip+one-hot",
                      "gpu_ms_per_batch_device": ms, "gpu_ms_per_batch_wall": wall, "samples_per_s": 2 / (wall / 1e3),
                      "algorithmic_bytes_per_batch": alg, "achieved_GBps": alg / ms / 1e6,
                      "cpu_oracle_s_per_sample": cpu_s, "cpu_threads": 1,
                      "speedup_vs_cpu_per_sample": cpu_s / (wall / 2 / 1e3), "resident_bytes": res.nbytes()}))


if __name__ == "__main__":
    main()
