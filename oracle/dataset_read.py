"""CPU restatement of the training-time read path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the unmodified reference
class ``data/dataset.py:propheseeTafDataset`` by ``oracle/make_golden.py`` (``tests/golden/
dataset_read.npz``).  Citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import numpy as np
import torch


def load_taf_volume(bins_half: bytes, bins_full: bytes, time_channels: int, img_size):
    """``data/dataset.py:294-308``: the two raw uint8 files of a sample -> float32
    ``[2K, H, W]`` (``bins{K/2}`` first, slot 0 = newest); K <= 4 reads one file only."""
    H, W = img_size
    if time_channels > 4:
        a = np.frombuffer(bins_half, dtype=np.uint8).reshape(int(time_channels), H, W).astype(np.float32)
        b = np.frombuffer(bins_full, dtype=np.uint8).reshape(int(time_channels), H, W).astype(np.float32)
        return np.concatenate([a, b], 0)
    return np.frombuffer(bins_half, dtype=np.uint8).reshape(int(time_channels * 2), H, W).astype(np.float32)


def augment_sample(volume: np.ndarray, input_img_size, sr: float, cx: int, cy: int, flip: bool) -> np.ndarray:
    """``data/dataset.py:219-234``: nearest resize to ``int(input * sr)``, trailing unit axes
    (``after_process`` :251-252), ``/ 255``, crop at ``(-cy, -cx)``, optional horizontal flip.
    Returns float32 ``[C, H_in, W_in, 1, 1]``."""
    Hin, Win = input_img_size
    img = torch.from_numpy(np.ascontiguousarray(volume))
    img = torch.nn.functional.interpolate(img[None, :, :, :], size=(int(Hin * sr), int(Win * sr)), mode="nearest")[0]
    img = img[:, :, :, None, None]
    img = img / 255
    img = img[:, -cy:Hin - cy, -cx:Win - cx]
    img = img.numpy()
    if flip:
        img = img[:, :, ::-1]
    return np.ascontiguousarray(img)


def augment_draws(input_img_size, sr: float, cx_draw: float, cy_draw: float):
    """The crop offsets the reference derives from its random draws (:153-161):
    ``int(uniform(int(W - sr W), 0))`` for ``sr > 1``, else 0."""
    if sr > 1.0:
        return int(cx_draw), int(cy_draw)
    return 0, 0
