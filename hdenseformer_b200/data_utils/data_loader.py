"""Normalisation / tensor-conversion transforms and the dataset wrapper with the reference's names and arguments
(/root/reference/data_utils/data_loader.py:8-210), executed on the GPU (see transformer_3d.py for the execution model).

`To_Tensor` is the stage that launches: it returns {'image': cuda float32 [channels, D, H, W], 'label': cuda float32
[num_class, D, H, W]} -- the reference returns the same tensors on the CPU and the trainer uploads them
(trainer.py:361-367).  `ResidentVolumes` keeps a whole training set in HBM (a HECKTOR-sized set of 500 two-channel 200^3
volumes is 32 GB of the B200's 180 GB), so an epoch reads nothing from the host.
Not built: CropResize (skimage.transform.resize with anti-aliasing; not in the shipped 3-D transform list, config.py:116)
and the 2-D transforms."""
import numpy as np
import torch

from ._plan import plan_of, pop_plan, _KEY


def hdf5_reader(data_path, key):
    """data_loader.py:8-13"""
    try:
        import h5py
    except ImportError as e:        # not in this image; the transforms below take numpy / torch arrays from any reader
        raise ImportError("hdf5_reader needs h5py") from e
    with h5py.File(data_path, 'r') as f:
        return np.asarray(f[key], dtype=np.float32)


def _no_plan(sample, name):
    if _KEY not in sample and isinstance(sample.get("image"), torch.Tensor) and sample["image"].is_cuda and "label" in sample \
            and isinstance(sample["label"], torch.Tensor) and sample["label"].dim() == 4:
        raise NotImplementedError(f"{name} after To_Tensor: place it before To_Tensor (the fused kernel normalises while it gathers)")


class Trunc_and_Normalize(object):
    """data_loader.py:16-36: truncate the gray scale to `scale` and map it to [0, 1]"""

    def __init__(self, scale=None):
        self.scale = scale
        if self.scale is not None:
            assert len(self.scale) == 2, 'scale error'

    def __call__(self, sample):
        _no_plan(sample, "Trunc_and_Normalize")
        plan_of(sample).normalise("trunc", self.scale[0], self.scale[1])
        return sample


class MRNormalize(object):
    """data_loader.py:39-50: every channel divided by its maximum (if non-zero), negatives set to 0"""

    def __call__(self, sample):
        _no_plan(sample, "MRNormalize")
        plan_of(sample).normalise("mr")
        return sample


class PETandCTNormalize(object):
    """data_loader.py:53-68: channel 0 (CT) clipped to mean +- w and scaled to [-1, 1]; channel 1 (PET) z-scored with the
    statistics of the (cropped) sample"""

    def __init__(self, mean=0, w=1024):
        self.mean = mean
        self.w = w

    def __call__(self, sample):
        _no_plan(sample, "PETandCTNormalize")
        plan_of(sample).normalise("petct", self.mean, self.w)
        return sample


class To_Tensor(object):
    """data_loader.py:126-159: image -> [channels, ...] float32, label -> one-hot [num_class, ...] float32 with channel 0 =
    background.  Runs the recorded chain on the GPU."""

    def __init__(self, num_class=2, input_channel=3, device=None):
        self.num_class = num_class
        self.channel = input_channel
        self.device = device

    def __call__(self, sample, out=None):
        plan = pop_plan(sample)
        image, label = plan.run(self.num_class, self.channel, device=self.device, out=out)
        return {'image': image, 'label': label}


class Compose(object):
    """torchvision.transforms.Compose (what trainer.py:233 wraps the list in)"""

    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, sample):
        for t in self.transforms:
            sample = t(sample)
        return sample


class ResidentVolumes(object):
    """Raw volumes uploaded once and kept in HBM; indexing gives the {'image', 'label'} sample dict the transforms take."""

    def __init__(self, samples, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.items = []
        for s in samples:
            img = torch.as_tensor(np.ascontiguousarray(s['image'], dtype=np.float32)) if not isinstance(s['image'], torch.Tensor) else s['image']
            lab = torch.as_tensor(np.ascontiguousarray(s['label'], dtype=np.float32)) if not isinstance(s['label'], torch.Tensor) else s['label']
            self.items.append((img.float().to(dev).contiguous(), lab.float().to(dev).contiguous()))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, index):
        img, lab = self.items[index]
        return {'image': img, 'label': lab}

    def nbytes(self):
        return sum(i.numel() * 4 + l.numel() * 4 for i, l in self.items)


class DataGenerator(torch.utils.data.Dataset):
    """data_loader.py:162-210.  `path_list` may also be a ResidentVolumes (or any sequence of sample dicts): then nothing
    is read from disk.  ROI selection (roi_number) is applied to the label like the reference."""

    def __init__(self, path_list, roi_number=None, num_class=2, transform=None, img_key='ct', lab_key='seg'):
        self.path_list = path_list
        self.roi_number = roi_number
        self.num_class = num_class
        self.transform = transform
        self.img_key = img_key
        self.lab_key = lab_key

    def __len__(self):
        return len(self.path_list)

    def _select_roi(self, label):
        """one ROI id -> binary mask; a list of ids -> classes 1..n in list order (everything else background)"""
        rois = self.roi_number if isinstance(self.roi_number, list) else [self.roi_number]
        assert self.num_class == len(rois) + 1
        is_t = isinstance(label, torch.Tensor)
        out = torch.zeros_like(label, dtype=torch.float32) if is_t else np.zeros(label.shape, dtype=np.float32)
        for cls, roi in enumerate(rois, start=1):
            out[label == roi] = cls
        return out

    def __getitem__(self, index):
        item = self.path_list[index]
        if isinstance(item, dict):
            image, label = item['image'], item['label']
        else:
            image = hdf5_reader(item, self.img_key)
            label = hdf5_reader(item, self.lab_key)
        if self.roi_number is not None:
            label = self._select_roi(label)
        sample = {'image': image, 'label': label}
        if self.transform is not None:
            sample = self.transform(sample)
        return sample


def collate_batch(dataset, indices, image_out=None, label_out=None):
    """Run the dataset's transform chain for `indices` into one batch ([B, M, D, H, W], [B, C, D, H, W]).  With preallocated
    outputs and a chain that ends in To_Tensor the kernels write straight into the batch tensors (no stacking copy)."""
    tf = dataset.transform
    direct = (image_out is not None and label_out is not None and isinstance(tf, Compose) and len(tf.transforms) > 0
              and isinstance(tf.transforms[-1], To_Tensor))
    if not direct:
        samples = [dataset[i] for i in indices]
        image, label = torch.stack([s['image'] for s in samples]), torch.stack([s['label'] for s in samples])
        if image_out is not None:
            image = image_out.copy_(image)
        if label_out is not None:
            label = label_out.copy_(label)
        return {'image': image, 'label': label}
    last, head = tf.transforms[-1], Compose(tf.transforms[:-1])
    dataset.transform = head
    try:
        for b, i in enumerate(indices):
            last(dataset[i], out=(image_out[b], label_out[b]))
    finally:
        dataset.transform = tf
    return {'image': image_out, 'label': label_out}
