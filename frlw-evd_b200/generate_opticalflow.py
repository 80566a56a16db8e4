"""Time-surface pair for optical flow -- the encoder half of the reference's
``generate_opticalflow.py``.

``generate_timesurface(events, volume1, volume2, end_stamp)`` keeps the reference signature
(:72-92).  The flow extraction that follows it in the reference (OpenCV TV-L1 on the two uint8
surfaces, :16-60) is not part of the event-representation path and is not provided here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def generate_timesurface(events, volume1, volume2, end_stamp=None):
    """``events``: float64 ``[N,4]`` (x, y, t, p), numpy (like the reference) or a CUDA tensor;
    ``volume1`` / ``volume2``: zero-initialised ``[H,W]`` arrays as at the reference's call site
    (:180-181) -- they only give the shape.  ``end_stamp`` is ignored, as in the reference, which
    overwrites it with the newest timestamp (:76).  Returns ``(volume1, volume2)`` float64, of the
    kind of ``events``; with no events the inputs come back unchanged (:75)."""
    if len(events) == 0:
        return volume1, volume2
    H, W = int(volume1.shape[0]), int(volume1.shape[1])
    if not torch.is_tensor(events):
        ev = np.asarray(events)
        x, y, t = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 2].astype(np.int64)
        keep = (x >= 0) & (y >= 0) & (x < 65536) & (y < 65536)          # uint16 columns; beyond the grid = dropped anyway
        soa = ops.EventStream.from_numpy(t[keep], x[keep], y[keep], np.zeros(int(keep.sum()), dtype=np.uint8))
        old, new = ops.timesurface(soa, (H, W))
        return old.cpu().numpy(), new.cpu().numpy()
    ev = events if events.is_cuda else events.cuda()
    x, y, t = ev[:, 0].to(torch.int64), ev[:, 1].to(torch.int64), ev[:, 2].to(torch.int64)
    keep = (x >= 0) & (y >= 0) & (x < 65536) & (y < 65536)
    soa = ops.EventStream(t[keep].to(torch.int32).view(torch.uint32), x[keep].to(torch.int16).view(torch.uint16),
                          y[keep].to(torch.int16).view(torch.uint16), torch.zeros(int(keep.sum()), dtype=torch.uint8, device=ev.device))
    return ops.timesurface(soa, (H, W))
