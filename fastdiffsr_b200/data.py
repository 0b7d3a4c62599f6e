"""Validation data path: the reference's folder-of-images LR/HR dataset (data/LRHR_dataset.py:10-128,
datatype 'img'; data/util.py:9-24 file discovery) with a batched, prefetching loader.

The reference's val loader yields one image per step from one worker (data/__init__.py:16-18) and
reads a *precomputed* bicubic image from `sr_{l}_{r}/`.  At thousands of images per second that
loader is the bottleneck, so here
  * items are decoded to uint8 (the reference's ToTensor + *2-1 happens on the device, bit-identically);
  * `sr_{l}_{r}/` is optional: the library's bit-exact PIL bicubic (`fdsr_bicubic_u8`) rebuilds it from
    `lr_{l}/`, and when the folder exists it is used as is, like the reference does;
  * batches are assembled in pinned memory by a background thread one batch ahead of the GPU.
"""
from __future__ import annotations

import os
import queue
import threading

import numpy as np
import torch

IMG_EXTENSIONS = ['.jpg', '.JPG', '.jpeg', '.JPEG', '.png', '.PNG', '.ppm', '.PPM', '.bmp', '.BMP', 'tif']


def is_image_file(filename):
    return any(filename.endswith(ext) for ext in IMG_EXTENSIONS)


def get_paths_from_images(path):
    """data/util.py:13-24."""
    assert os.path.isdir(path), '{:s} is not a valid directory'.format(path)
    images = []
    for dirpath, _, fnames in sorted(os.walk(path)):
        for fname in sorted(fnames):
            if is_image_file(fname):
                images.append(os.path.join(dirpath, fname))
    assert images, '{:s} has no valid image file'.format(path)
    return sorted(images)


def _load_rgb_u8(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.array(im.convert("RGB"), dtype=np.uint8)  # (copy: PIL buffers are read-only)


def u8_to_tensor(img_u8: torch.Tensor):
    """(B,H,W,3) uint8 -> (B,3,H,W) fp32 in [-1,1]: torchvision ToTensor (/255) then *2-1
    (data/util.py:66-75 with min_max=(-1,1)), same operation order."""
    return (img_u8.permute(0, 3, 1, 2).float() / 255.0) * 2.0 - 1.0


class LRHRDataset:
    """Folder layout of the reference: `{dataroot}/hr_{r}`, `{dataroot}/lr_{l}`, optional `{dataroot}/sr_{l}_{r}`."""

    def __init__(self, dataroot, datatype="img", l_resolution=64, r_resolution=256, split="val", data_len=-1,
                 need_LR=True, img_mask="no"):
        if datatype != "img":
            raise NotImplementedError("data_type [{:s}] is not recognized (lmdb datasets are a training-side "
                                      "format of the reference; convert to image folders)".format(str(datatype)))
        if split == "train":
            raise NotImplementedError("the training data path (random flips, lmdb) is outside the B200 sampling path")
        self.l_res, self.r_res, self.split = l_resolution, r_resolution, split
        self.hr_path = get_paths_from_images('{}/hr_{}'.format(dataroot, r_resolution))
        sr_dir = '{}/sr_{}_{}'.format(dataroot, l_resolution, r_resolution)
        lr_dir = '{}/lr_{}'.format(dataroot, l_resolution)
        self.sr_path = get_paths_from_images(sr_dir) if os.path.isdir(sr_dir) else None
        self.lr_path = get_paths_from_images(lr_dir) if os.path.isdir(lr_dir) else None
        if self.sr_path is None and self.lr_path is None:
            raise FileNotFoundError(f"{dataroot}: neither {os.path.basename(sr_dir)}/ nor {os.path.basename(lr_dir)}/ exists")
        for p in (self.sr_path, self.lr_path):
            if p is not None and len(p) != len(self.hr_path):
                raise ValueError(f"{dataroot}: folder sizes differ ({len(p)} vs {len(self.hr_path)} HR images)")
        self.dataset_len = len(self.hr_path)
        self.data_len = self.dataset_len if (data_len is None or data_len <= 0) else min(data_len, self.dataset_len)

    def __len__(self):
        return self.data_len

    def get_u8(self, index):
        """uint8 HWC arrays: always 'HR'; 'LR' and/or 'SR' as present on disk."""
        item = {"HR": _load_rgb_u8(self.hr_path[index]), "Index": index, "path": self.hr_path[index]}
        if self.lr_path is not None:
            item["LR"] = _load_rgb_u8(self.lr_path[index])
        if self.sr_path is not None:
            item["SR"] = _load_rgb_u8(self.sr_path[index])
        return item

    def __getitem__(self, index):
        """The reference's item (fp32 CHW tensors in [-1,1]; 'SR' needs the sr_ folder)."""
        it = self.get_u8(index)
        out = {"Index": index}
        for k in ("HR", "SR", "LR"):
            if k in it:
                out[k] = u8_to_tensor(torch.from_numpy(it[k])[None])[0]
        return out


def create_dataset(dataset_opt, phase):
    """data/__init__.py:24-40."""
    return LRHRDataset(dataroot=dataset_opt['dataroot'], datatype=dataset_opt['datatype'],
                       l_resolution=dataset_opt['l_resolution'], r_resolution=dataset_opt['r_resolution'],
                       split=phase, data_len=dataset_opt['data_len'] if dataset_opt['data_len'] is not None else -1,
                       need_LR=(dataset_opt['mode'] == 'LRHR'))


class BatchLoader:
    """Iterates `indices` of a dataset in batches of uint8 arrays stacked in pinned host memory, decoding one
    batch ahead in a background thread.  Yields dicts {'HR','LR'?,'SR'?: pinned uint8 (B,H,W,3), 'Index': list}."""

    def __init__(self, dataset, batch_size, indices=None, prefetch=2, pin=True):
        self.ds = dataset
        self.bs = max(1, int(batch_size))
        self.indices = list(range(len(dataset))) if indices is None else list(indices)
        self.prefetch = prefetch
        self.pin = pin and torch.cuda.is_available()

    def __len__(self):
        return (len(self.indices) + self.bs - 1) // self.bs

    def _assemble(self, idxs):
        items = [self.ds.get_u8(i) for i in idxs]
        out = {"Index": list(idxs), "path": [it["path"] for it in items]}
        for k in ("HR", "LR", "SR"):
            if k in items[0]:
                arr = torch.from_numpy(np.stack([it[k] for it in items], 0))
                out[k] = arr.pin_memory() if self.pin else arr
        return out

    def __iter__(self):
        q = queue.Queue(maxsize=self.prefetch)
        chunks = [self.indices[i:i + self.bs] for i in range(0, len(self.indices), self.bs)]

        def work():
            try:
                for ch in chunks:
                    q.put(self._assemble(ch))
                q.put(None)
            except BaseException as e:  # surface decode errors in the consumer
                q.put(e)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            yield item
