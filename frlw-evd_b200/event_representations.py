"""``event_representations`` -- drop-in for the reference's pybind11 extension
(``data/event_representation_tool/src/event_queue_tensor.cpp:120-123``).

``event_queue_tensor(events, queue_length, B, H, W, start_times, event_window_abin)`` takes
and returns numpy arrays like the extension: ``events`` float32 ``[N,6]`` rows
(b, x, y, t, p, z), ``start_times`` int32 ``[B]``; result float64 ``[2, Q, 2, B, H, W]``.
The computation runs on the current CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream


def event_queue_tensor(events, queue_length, B, H, W, start_times, event_window_abin):
    if not torch.cuda.is_available():
        raise _lib.EvrepError("event_queue_tensor needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    ev = torch.from_numpy(np.ascontiguousarray(events, dtype=np.float32).reshape(-1, 6)).to(dev)
    start = torch.from_numpy(np.ascontiguousarray(start_times, dtype=np.int32)).to(dev)
    Q = int(queue_length)
    totals = torch.empty(2 * B * H * W, dtype=torch.float32, device=dev)
    out = torch.empty((2, Q, 2, B, H, W), dtype=torch.float64, device=dev)
    _lib.call("evrep_event_queue_tensor", _ptr(ev), ev.shape[0], Q, B, H, W, _ptr(start), int(event_window_abin),
              _ptr(totals), _ptr(out), _stream(dev))
    return out.cpu().numpy()
