"""Shared pieces of the four ``generate_*`` command-line drivers: dataset geometry, the
``train/val/test`` directory walk, GPU residency of a recording and the raw-uint8 writer.

Layout written (reader contract ``data/dataset.py:241-249,294-308`` of the reference):
headerless C-order ``uint8`` tensors in files named ``<recording>_<label t>.npy``.
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops
from .io import PSEELoader, npy_events_tools


@dataclass
class Geometry:
    """Sensor / target shapes and the resize policy R1 (``generate_taf.py:93-104,216-222``):
    a target smaller than the sensor (gen4) is encoded directly on the target grid through
    truncated float64 coordinate scaling; otherwise (gen1) the encoder runs on the sensor
    grid and the tensor is nearest-resized."""
    shape: tuple
    target: tuple
    device: str = "cuda"
    coord_maps: Optional[ops.CoordMaps] = None
    resize_maps: Optional[tuple] = None

    @classmethod
    def for_dataset(cls, dataset: str, device="cuda") -> "Geometry":
        if dataset == "gen4":
            g = cls((720, 1280), (512, 640), device)
        else:
            g = cls((240, 304), (256, 320), device)
        if g.downscale:
            g.coord_maps = ops.make_coord_maps(g.shape, g.target, device)
        else:
            g.resize_maps = ops.nearest_maps(g.shape, g.target, device)
        return g

    @property
    def downscale(self) -> bool:
        return self.target[0] < self.shape[0]

    @property
    def grid(self) -> tuple:
        """The grid the encoder kernels run on."""
        return self.target if self.downscale else self.shape

    def to_target(self, volume: torch.Tensor) -> torch.Tensor:
        return volume if self.downscale else ops.nearest_resize(volume, self.target, self.resize_maps)


def parse_args(default_dataset: str, argv=None):
    parser = argparse.ArgumentParser(description="event-representation generator (B200)")
    parser.add_argument("-raw_dir", type=str)      # "train, val, test" level directory of the event files
    parser.add_argument("-label_dir", type=str)    # "train, val, test" level directory of the annotations
    parser.add_argument("-target_dir", type=str)   # output directory
    parser.add_argument("-dataset", type=str, default=default_dataset)   # gen1 / gen4
    return parser.parse_args(argv)


def iter_recordings(raw_dir, label_dir):
    """Yield ``(mode, name, event file, label timestamps)`` like the reference's loops
    (``generate_taf.py:112-151``): unreadable modes are skipped silently."""
    for mode in ("train", "val", "test"):
        try:
            listing = os.listdir(os.path.join(raw_dir, mode))
        except Exception:
            continue
        for name in [f[:-7] for f in listing if f[-3:] == "dat"]:
            labels = npy_events_tools.read_label_times(os.path.join(label_dir, mode, name + "_bbox.npy"))
            yield mode, name, os.path.join(raw_dir, mode, name + "_td.dat"), labels


class DeviceRecording:
    """A recording staged for the GPU: the raw payload sits in pinned host memory
    (``raw_pinned``); with ``decode=True`` it is also copied to the device and decoded by the
    CUDA decoder into SoA buffers (``events``).  The memory-mapped loader answers the seek
    queries."""

    def __init__(self, path: str, device="cuda", decode=True):
        self.loader = PSEELoader(path)
        payload = self.loader.raw_bytes()
        self.raw_pinned = torch.empty(payload.shape[0], dtype=torch.uint8, pin_memory=payload.shape[0] > 0)
        np.copyto(self.raw_pinned.numpy(), payload)      # file (page cache) -> pinned host memory, once
        self.n_events = payload.shape[0] // 8
        self.events = ops.decode_dat(self.raw_pinned.to(device, non_blocking=True)) if decode else None


def dump_u8(tensor_u8: torch.Tensor, *path) -> None:
    os.makedirs(os.path.join(*path[:-1]), exist_ok=True)
    tensor_u8.cpu().numpy().tofile(os.path.join(*path))


class AsyncWriter:
    """Writes raw uint8 arrays to files on a small thread pool (``ndarray.tofile`` releases
    the GIL), so the file system works while the GPU encodes the next recording."""

    def __init__(self, workers: int = 4):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._pending = []
        self._dirs = set()
        self.bytes_written = 0

    def put(self, array: np.ndarray, *path):
        """Queue one file; returns its future (the array must stay untouched until it is done)."""
        folder = os.path.join(*path[:-1])
        if folder not in self._dirs:
            os.makedirs(folder, exist_ok=True)
            self._dirs.add(folder)
        self.bytes_written += array.nbytes
        future = self._pool.submit(array.tofile, os.path.join(*path))
        self._pending.append(future)
        return future

    def drain(self) -> None:
        """Block until everything submitted so far is on disk (buffers may then be reused)."""
        for f in self._pending:
            f.result()
        self._pending = []

    def close(self) -> None:
        self.drain()
        self._pool.shutdown()


class PinnedRing:
    """Device -> host -> file hand-over of the drivers' output chunks without a blocking copy on the
    encode loop and with a bounded amount of page-locked memory: a ring of pinned buffers, a copy
    stream, and the ``AsyncWriter``.  ``push(dev_u8, emit)`` queues the device->host copy of a chunk
    behind the kernels that produce it and returns at once; ``emit(host_array)`` -- which calls
    ``writer.put`` for every file of the chunk and returns the futures -- runs when the NEXT chunk is
    pushed (or at ``flush``), i.e. while the GPU is already encoding again.  A buffer is reused only
    after its files are on disk."""

    def __init__(self, device="cuda", slots: int = 3):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [dict(buf=None, done=None, emit=None, futures=[], shape=None) for _ in range(slots)]
        self.count = 0
        self.pinned_bytes = 0

    def _retire(self, slot):
        if slot["emit"] is not None:
            slot["done"].synchronize()
            n = int(np.prod(slot["shape"]))
            slot["futures"] = list(slot["emit"](slot["buf"][:n].numpy().reshape(slot["shape"])) or [])
            slot["emit"] = None

    def push(self, dev_u8: torch.Tensor, emit) -> None:
        assert dev_u8.dtype == torch.uint8 and dev_u8.is_contiguous()
        if self.count:
            self._retire(self.slots[(self.count - 1) % len(self.slots)])      # the previous chunk has had a whole encode to land
        slot = self.slots[self.count % len(self.slots)]
        self._retire(slot)
        for f in slot["futures"]:
            f.result()                                                        # its files are written: the buffer is free
        slot["futures"] = []
        n = dev_u8.numel()
        if slot["buf"] is None or slot["buf"].numel() < n:
            self.pinned_bytes += n - (0 if slot["buf"] is None else slot["buf"].numel())
            slot["buf"] = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.copy_stream.wait_event(ready)
        with torch.cuda.stream(self.copy_stream):
            slot["buf"][:n].copy_(dev_u8.view(-1), non_blocking=True)
            dev_u8.record_stream(self.copy_stream)
            slot["done"] = torch.cuda.Event()
            slot["done"].record(self.copy_stream)
        slot["emit"], slot["shape"] = emit, tuple(dev_u8.shape)
        self.count += 1

    def flush(self) -> None:
        """Hand every chunk still in flight to the writer and wait for the files."""
        for k in range(len(self.slots)):
            slot = self.slots[(self.count + k) % len(self.slots)]
            self._retire(slot)
        for slot in self.slots:
            for f in slot["futures"]:
                f.result()
            slot["futures"] = []
