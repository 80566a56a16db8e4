"""Training-time read path on the GPU.

The reference decodes every sample on a CPU worker (``data/dataset.py:219-234``): raw uint8
file(s) -> float32, nearest resize by the augmentation scale, ``/ 255``, crop, flip.  Here the
raw bytes go to the device as they are (one byte per value instead of four) and one kernel
(``evrep_load_samples``) does the rest for a whole batch.  The augmentation *parameters* stay
with the caller -- they are coupled to the box transforms of ``__getitem__`` (:140-208) -- and
are passed in as ``(sr, cx, cy, flip)`` exactly as the reference draws them.
"""
from __future__ import annotations

import ctypes
import os
from typing import Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from ..ops import _need_cuda, _ptr, _stream


def taf_file_names(data_dir: str, mode: str, file_name: str, timestamp: int, time_channels: int):
    """Paths of the file(s) of one sample (``data/dataset.py:294-308``): ``bins{K/2}`` and
    ``bins{K}`` for K > 4, ``bins{K}`` alone otherwise."""
    root = os.path.join(data_dir, mode)
    name = file_name + "_" + str(timestamp) + ".npy"
    if time_channels > 4:
        return [os.path.join(root, "bins{0}".format(int(time_channels // 2)), name),
                os.path.join(root, "bins{0}".format(int(time_channels)), name)]
    return [os.path.join(root, "bins{0}".format(int(time_channels)), name)]


def read_sample_bytes(paths: Sequence[str], out: torch.Tensor = None) -> torch.Tensor:
    """Concatenated raw bytes of a sample's files as a (pinned) uint8 host tensor."""
    sizes = [os.path.getsize(p) for p in paths]
    if out is None:
        out = torch.empty(sum(sizes), dtype=torch.uint8).pin_memory() if torch.cuda.is_available() \
            else torch.empty(sum(sizes), dtype=torch.uint8)
    view, at = out.numpy(), 0
    for p, n in zip(paths, sizes):
        with open(p, "rb") as fh:
            got = fh.readinto(memoryview(view[at:at + n]))
        assert got == n, p
        at += n
    return out


def augment_batch(volumes: torch.Tensor, input_img_size, params: Sequence[Tuple[float, int, int, bool]],
                  out: torch.Tensor = None) -> torch.Tensor:
    """``volumes``: uint8 CUDA tensor ``[n, C, Hs, Ws]`` (file contents); ``params``: per sample
    ``(sr, cx, cy, flip)`` as drawn by the reference (:141-161).  Returns float32
    ``[n, C, H_in, W_in]``; ``out[i][..., None, None]`` is the reference's image of sample i."""
    _need_cuda(volumes)
    assert volumes.dtype == torch.uint8 and volumes.is_contiguous() and volumes.dim() == 4
    n, C, Hs, Ws = volumes.shape
    Hin, Win = int(input_img_size[0]), int(input_img_size[1])
    assert len(params) == n
    desc = np.empty((n, 5), dtype=np.int32)
    for i, (sr, cx, cy, flip) in enumerate(params):
        up_h, up_w = int(Hin * sr), int(Win * sr)                    # F.interpolate(size=...) at :223
        if not (cy <= 0 and cx <= 0 and Hin - cy <= up_h and Win - cx <= up_w):
            raise ValueError("crop (%d, %d) leaves the resized image %dx%d" % (cy, cx, up_h, up_w))
        desc[i] = (up_h, up_w, cy, cx, 1 if flip else 0)
    aug = torch.from_numpy(desc).to(volumes.device)
    if out is None:
        out = torch.empty((n, C, Hin, Win), dtype=torch.float32, device=volumes.device)
    _lib.call("evrep_load_samples", _ptr(volumes), C * Hs * Ws, n, C, Hs, Ws, _ptr(aug), Hin, Win, _ptr(out),
              _stream(volumes.device))
    return out


def load_sample(paths: Sequence[str], channels: int, img_size, input_img_size, sr=1.0, cx=0, cy=0, flip=False,
                device="cuda") -> torch.Tensor:
    """One sample end to end, shaped like the reference's ``img``: float32 ``[C, H_in, W_in, 1, 1]``."""
    raw = read_sample_bytes(paths).to(device, non_blocking=True).view(1, channels, int(img_size[0]), int(img_size[1]))
    return augment_batch(raw, input_img_size, [(sr, cx, cy, flip)])[0][:, :, :, None, None]
