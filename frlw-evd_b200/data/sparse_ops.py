"""Batched online encoders -- drop-in for the reference's ``data/sparse_ops.py``.

Same six functions, same uniform plugin signature as called by ``data/fetcher.py:53``:
``f(events, B, shape, iter, memory_or_past, events_window, volume_bins, infer_time) ->
(volume, memory)``.  ``events`` is a CUDA tensor ``[N,5]`` with columns (b, x, y, t, p)
(``[N,7]`` = (b, x, y, t, c, p, feature) for ``generate_taf_cuda``).
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from ..ops import _need_cuda, _ptr, _stream


def _f64(events):
    _need_cuda(events)
    return events.to(torch.float64).contiguous()


def _splat(events, B, H, W, C, mode, K, window, iter_, infer):
    acc = torch.empty((B * H * W, C, 2), dtype=torch.float32, device=events.device)
    _lib.call("evrep_sparse_splat", _ptr(events), events.shape[0], B, H, W, C, mode, float(K), float(window),
              float(iter_), float(infer), _ptr(acc), _stream(events.device))
    return acc


def _planar(img, B, H, W):
    C2 = img.shape[1] * 2
    out = torch.empty((B, C2, H, W, 1), dtype=torch.float32, device=img.device)
    _lib.call("evrep_pixel_major_to_planar", _ptr(img), B, H * W, C2, _ptr(out), _stream(img.device))
    return out


def generate_agile_event_volume_cuda(events, B, shape, iter, past_volume=None, events_window=50000,
                                     volume_bins=5, infer_time=10000):
    """``data/sparse_ops.py:4-35``.  Returns ``(f32 [B,2K,H,W,1], state f32 [BHW,K,2,1])``.
    In incremental mode ``past_volume[:, -1]`` is updated IN PLACE like the reference does."""
    H, W = shape
    events = _f64(events)
    if past_volume is None:
        img = _splat(events, B, H, W, volume_bins, 0, volume_bins, events_window, 0, 0)
    else:
        fresh = _splat(events, B, H, W, 2, 1, volume_bins, events_window, iter, infer_time)
        assert past_volume.is_cuda and past_volume.is_contiguous() and past_volume.dtype == torch.float32
        img = torch.empty((B * H * W, volume_bins, 2), dtype=torch.float32, device=events.device)
        _lib.call("evrep_sparse_agile_shift", _ptr(past_volume), _ptr(fresh), B * H * W, volume_bins, _ptr(img),
                  _stream(events.device))
    return _planar(img, B, H, W), img.view(B * H * W, volume_bins, 2, 1)


def generate_event_volume_cuda(events, B, shape, iter, memory=None, events_window=50000, volume_bins=5,
                               infer_time=10000):
    """``data/sparse_ops.py:37-69``: raw-event memory (tensor plumbing: concatenate + time
    filter), then the splat with ``t* = (K-1) t / window``."""
    H, W = shape
    _need_cuda(events)
    if memory is not None:
        events = torch.cat([memory, events])
    memory = events[events[:, 3] >= iter - events_window + infer_time]
    img = _splat(_f64(events), B, H, W, volume_bins, 2, volume_bins, events_window, 0, 0)
    return _planar(img, B, H, W), memory


def generate_taf_cuda(events, B, shape, iter, past_volume=None, events_window=50000, volume_bins=5,
                      infer_time=10000):
    """``data/sparse_ops.py:72-85``.  Returns ``(f32 [B,2K,H,W,2], None)``."""
    H, W = shape
    events = _f64(events)
    C = volume_bins * 2
    out = torch.empty((B, C, H, W, 2), dtype=torch.float32, device=events.device)
    _lib.call("evrep_sparse_taf", _ptr(events), events.shape[0], B, H, W, C, _ptr(out), _stream(events.device))
    return out, None


def generate_event_frame_cuda(events, B, shape, iter, past_volume=None, events_window=50000, volume_bins=5,
                              infer_time=10000):
    """``data/sparse_ops.py:88-107``.  Returns ``(f32 [B,2,H,W,1], None)``."""
    H, W = shape
    events = _f64(events)
    out = torch.empty((B, 2, H, W, 1), dtype=torch.float32, device=events.device)
    _lib.call("evrep_sparse_event_frame", _ptr(events), events.shape[0], B, H, W, _ptr(out), _stream(events.device))
    return out, None


def sparseToDense(locations, features, shape):
    """``data/sparse_ops.py:109-121``: ``locations`` ``[N,3]`` (b, y, x), ``features`` ``[N,C]``."""
    B, H, W = shape
    _need_cuda(locations, features)
    loc = locations.to(torch.int64).contiguous()
    feat = features.to(torch.float32).contiguous()
    C = feat.shape[-1]
    out = torch.empty((B, H, W, C), dtype=torch.float32, device=feat.device)
    _lib.call("evrep_sparse_to_dense", _ptr(loc), _ptr(feat), loc.shape[0], B, H, W, C, _ptr(out), _stream(feat.device))
    return out


def denseToSparse(dense_tensor):
    """``data/sparse_ops.py:123-135``: rows with a non-zero |.|-sum.  Index bookkeeping only
    (``nonzero`` + gather on the device); locations are (y, x, b)."""
    _need_cuda(dense_tensor)
    nz = torch.nonzero(torch.abs(dense_tensor).sum(dim=-1))
    locations = torch.cat((nz[:, 1:], nz[:, 0, None]), dim=-1)
    features = dense_tensor[nz[:, 0], nz[:, 1], nz[:, 2]]
    return locations, features
