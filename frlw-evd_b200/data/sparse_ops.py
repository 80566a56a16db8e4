"""Batched online encoders -- drop-in for the reference's ``data/sparse_ops.py``.

Same six functions (plus ``generate_taf_online_cuda``, a streaming TAF with the state carried on the
device), same uniform plugin signature as called by ``data/fetcher.py:53``:
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


def _compact_scratch(rows, device):
    need = _lib.load().evrep_compact_scratch_bytes(int(rows))
    return torch.empty(max(int(need) // 4, 2), dtype=torch.int32, device=device)


def _kept(count_ptr, scratch):
    """Number of rows a compaction kept: one 4-byte read back (what ``torch.nonzero`` does as well)."""
    offset = (count_ptr.value - scratch.data_ptr()) // 4
    return int(scratch[offset].item())


def generate_event_volume_cuda(events, B, shape, iter, memory=None, events_window=50000, volume_bins=5,
                               infer_time=10000):
    """``data/sparse_ops.py:37-69``: the raw-event memory (concatenate, keep ``t >= iter - window +
    infer_time``: one copy kernel + a stable compaction), then the splat with ``t* = (K-1) t / window``."""
    H, W = shape
    new = _f64(events)
    old = None if memory is None else _f64(memory)
    n_old = 0 if old is None else old.shape[0]
    n = n_old + new.shape[0]
    merged = torch.empty((n, 5), dtype=torch.float64, device=new.device)
    kept = torch.empty((n, 5), dtype=torch.float64, device=new.device)
    scratch = _compact_scratch(n, new.device)
    count = ctypes.c_void_p(0)
    _lib.call("evrep_event_memory_update", _ptr(old), n_old, _ptr(new), new.shape[0], float(iter - events_window + infer_time),
              _ptr(merged), _ptr(kept), _ptr(scratch), ctypes.byref(count), _stream(new.device))
    img = _splat(merged, B, H, W, volume_bins, 2, volume_bins, events_window, 0, 0)
    memory = kept[:_kept(count, scratch)] if n else kept
    return _planar(img, B, H, W), memory.to(events.dtype)


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
    """``data/sparse_ops.py:123-135``: rows with a non-zero |.|-sum, in row-major order (a flag pass, a
    scan and a scatter on the device); locations are (spatial indices..., batch)."""
    _need_cuda(dense_tensor)
    dense = dense_tensor.to(torch.float32).contiguous()
    lead, C = tuple(dense.shape[:-1]), int(dense.shape[-1])
    assert 1 <= len(lead) <= 4, "denseToSparse handles batch + up to three spatial dimensions"
    rows = 1
    for v in lead:
        rows *= int(v)
    locations = torch.empty((rows, len(lead)), dtype=torch.int64, device=dense.device)
    features = torch.empty((rows, C), dtype=torch.float32, device=dense.device)
    scratch = _compact_scratch(rows, dense.device)
    sizes = (ctypes.c_int64 * len(lead))(*[int(v) for v in lead])
    count = ctypes.c_void_p(0)
    _lib.call("evrep_dense_to_sparse", _ptr(dense), ctypes.cast(sizes, ctypes.c_void_p), len(lead), C, _ptr(locations),
              _ptr(features), _ptr(scratch), ctypes.byref(count), _stream(dense.device))
    n = _kept(count, scratch) if rows else 0
    return locations[:n], features[:n].to(dense_tensor.dtype)


def generate_taf_online_cuda(events, B, shape, iter, memory=None, events_window=50000, volume_bins=8, infer_time=10000):
    """Online Temporal Active Focus behind the same plugin signature (no counterpart in the reference's
    ``data/sparse_ops.py``; it is ``generate_taf.py:195-222`` run step by step): the FIFO state of the B
    recordings of the batch stays on the device in ``memory`` and every ``fetch()`` step pushes its
    ``infer_time`` bins (``events_window / infer_time`` of them on the first step, one afterwards).
    Nothing here synchronises.  Returns ``(f32 [B,2K,H,W,1], memory)``."""
    H, W = shape
    ev = _f64(events)
    K = int(volume_bins)
    if memory is None:
        state = torch.full((B, H, W, 2, K), -6000.0, dtype=torch.float32, device=ev.device)       # generate_taf.py:207-209
        need = _lib.load().evrep_taf_online_scratch_bytes(B, H, W)
        memory = {"state": state, "scratch": torch.zeros(int(need), dtype=torch.uint8, device=ev.device)}
        start = iter - events_window
    else:
        start = iter - infer_time
    n_bins = max(1, -(-int(iter - start) // int(infer_time)))
    out = torch.empty((B, 2 * K, H, W, 1), dtype=torch.float32, device=ev.device)
    for k in range(n_bins):
        lo = start + k * infer_time
        last = k == n_bins - 1
        _lib.call("evrep_taf_online_bin", _ptr(ev), ev.shape[0], float(lo), float(iter if last else lo + infer_time),
                  float(infer_time + 1e-8), B, H, W, K, _ptr(memory["state"]), _ptr(out) if last else _ptr(None),
                  _ptr(memory["scratch"]), _stream(ev.device))
    return out, memory
