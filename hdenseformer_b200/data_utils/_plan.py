"""Deferred execution plan behind the GPU transforms: every transform records its parameters, To_Tensor launches."""
import numpy as np
import torch

from .. import ops

_KEY = "_hdf_plan"
_STAGES = {"crop": 1, "normalise": 2, "warp": 3, "flip": 4}


class Plan:
    def __init__(self, sample):
        image = sample["image"]
        self.image, self.label = image, sample.get("label")
        shape = tuple(image.shape)
        if len(shape) == 4:
            self.M, self.size = shape[0], shape[1:]
        elif len(shape) == 3:
            self.M, self.size = 1, shape
        else:
            raise ValueError(f"image must be [M, D, H, W] or [D, H, W], got {shape}")
        self.origin = (0, 0, 0)
        self.norm, self.p0, self.p1 = None, 0.0, 0.0
        self.affine, self.warp_classes = None, None
        self.flip_axis = 0
        self.stage = 0
        self.done = set()

    def _enter(self, name):
        st = _STAGES[name]
        if st < self.stage or name in self.done:
            raise NotImplementedError(
                f"GPU input pipeline: '{name}' after stage {self.stage} -- the fused kernel runs the reference's list order "
                "crop -> normalise -> warp -> flip -> To_Tensor (trainer.py:128-141), each at most once")
        self.stage = st
        self.done.add(name)

    def crop(self, origin, size):
        self._enter("crop")
        self.origin, self.size = tuple(int(o) for o in origin), tuple(int(s) for s in size)

    def normalise(self, mode, p0=0.0, p1=0.0):
        self._enter("normalise")
        self.norm, self.p0, self.p1 = mode, float(p0), float(p1)

    def warp(self, warp_mat, num_class):
        self._enter("warp")
        self.affine = np.ascontiguousarray(np.asarray(warp_mat, dtype=np.float64)[:3, :4])
        self.warp_classes = int(num_class)

    def flip(self, axis):
        self._enter("flip")
        self.flip_axis = int(axis)

    # ------------------------------------------------------------------ launch
    @staticmethod
    def _to_device(a, dev):
        if a is None:
            return None
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        if t.dtype != torch.float32:
            t = t.float()
        return t.to(dev, non_blocking=True).contiguous()

    def run(self, num_class, input_channel, device=None, out=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ops._C.HDFError("the GPU input pipeline has no CPU path")
        vol = self._to_device(self.image, dev)
        if vol.dim() == 3:
            vol = vol.unsqueeze(0)
        lab = self._to_device(self.label, dev)
        if self.warp_classes is not None and self.warp_classes != num_class:
            raise ValueError(f"RandomTranslationRotationZoom3D(num_class={self.warp_classes}) vs To_Tensor(num_class={num_class})")
        D, H, W = self.size
        if out is None:
            img_out = torch.empty((self.M, D, H, W), dtype=torch.float32, device=dev)
            lab_out = torch.empty((num_class, D, H, W), dtype=torch.float32, device=dev) if lab is not None else None
        else:
            img_out, lab_out = out
        aff = None if self.affine is None else torch.from_numpy(self.affine).to(dev, non_blocking=True)
        ops.prep_sample(vol, lab, self.origin, self.size, self.norm, self.p0, self.p1, aff, self.flip_axis, num_class, img_out, lab_out)
        # data_loader.py:138-141: multi-channel -> the first `input_channel` channels; single channel -> a leading axis
        image = img_out[:input_channel] if input_channel > 1 else img_out[:1]
        return image, lab_out


def plan_of(sample) -> Plan:
    p = sample.get(_KEY)
    if p is None:
        p = Plan(sample)
        sample[_KEY] = p
    return p


def pop_plan(sample) -> Plan:
    return plan_of(sample) if _KEY not in sample else sample.pop(_KEY)
