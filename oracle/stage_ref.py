"""Stages the UNMODIFIED reference modules of the hot path into the git-ignored `oracle/_ref/` (test / bench
infrastructure only, never imported by the product package).

The reference is plain Python with no build or install target (no setup.py / pyproject), so "building the reference
where it compiles" reduces to making its three hot-path files importable on the GPU box, where /root/reference does not
exist: `oracle/_ref/` travels with gpurun snapshots but stays out of git history.  Staged files (verbatim copies):
    models/HDenseFormer.py, models/HDenseFormer_2D.py, loss/{combine_loss,dice_loss,cross_entropy}.py (+ empty __init__)
`load()` imports them as the top-level packages `models` / `loss` exactly as the reference's trainer does
(trainer.py:642-647, 763-765)."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/HDenseFormer.py", "models/HDenseFormer_2D.py", "loss/__init__.py",
         "loss/combine_loss.py", "loss/dice_loss.py", "loss/cross_entropy.py"]


def stage(src: str = "/root/reference") -> bool:
    """Copy the hot-path modules from `src` if it exists (build container); returns True if `_ref` is usable."""
    if os.path.isdir(src):
        manifest = {}
        for f in FILES:
            s, d = os.path.join(src, f), os.path.join(DST, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[f] = hashlib.sha256(open(d, "rb").read()).hexdigest()
        json.dump({"source": src, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def load():
    """Returns (HDenseFormer_32, HDenseFormer_16, CEPlusDice, DeepSuperloss) of the staged reference, or None."""
    if not available():
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from models.HDenseFormer import HDenseFormer_16, HDenseFormer_32     # noqa: E402  (the reference's own package names)
    from loss.combine_loss import CEPlusDice, DeepSuperloss              # noqa: E402
    return HDenseFormer_32, HDenseFormer_16, CEPlusDice, DeepSuperloss


if __name__ == "__main__":
    print("staged" if stage() else "reference not available")
