"""Torch-tensor front end of the C ABI: device buffers, scratch and launches.

PyTorch is used for device memory, streams and (elsewhere) ``torch.distributed`` only;
all arithmetic happens in ``libevrep.so``.  Every function here requires CUDA tensors and
raises if the library or a device is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

COORD_LUT_LEN = 16384
TAF_INIT = -6000.0


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _records(rows, struct):
    """A list of tuples as a packed array of the ctypes ``struct`` (one numpy conversion instead of a Python object
    per row).  Returns ``(keep-alive array, c_void_p)``."""
    dtype = np.dtype([(name, np.dtype(ctype)) for name, ctype in struct._fields_], align=False)
    assert dtype.itemsize == ctypes.sizeof(struct), struct
    n = len(rows)
    arr = np.zeros(max(n, 1), dtype=dtype)
    if n:
        cols = np.asarray(rows, dtype=np.int64).reshape(n, len(struct._fields_))
        for k, (name, _) in enumerate(struct._fields_):
            arr[name][:n] = cols[:, k]
    return arr, ctypes.c_void_p(arr.ctypes.data)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.EvrepError("libevrep encoders need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)


# ------------------------------------------------------------------------ geometry
@dataclass
class CoordMaps:
    """Device LUTs for the gen4 down-scaling policy (``generate_taf.py:103-104,216-218``)."""
    xmap: torch.Tensor      # u16 [16384]
    ymap: torch.Tensor
    sensor_shape: tuple = (COORD_LUT_LEN, COORD_LUT_LEN)   # raw (H, W) range the maps cover


def make_coord_maps(sensor_shape, target_shape, device) -> CoordMaps:
    """``coord * (target / sensor)`` in float64, truncated; coordinates outside the sensor
    map to 0xFFFF (dropped)."""
    (H, W), (Ht, Wt) = sensor_shape, target_shape
    rh, rw = Ht / H, Wt / W

    def lut(n_in, ratio):
        full = np.full(COORD_LUT_LEN, 0xFFFF, dtype=np.uint16)
        full[:n_in] = (np.arange(n_in, dtype=np.float64) * ratio).astype(np.int64).astype(np.uint16)
        return torch.from_numpy(full).to(device)

    return CoordMaps(xmap=lut(W, rw), ymap=lut(H, rh), sensor_shape=(H, W))


def nearest_maps(in_shape, out_shape, device):
    """Legacy ``F.interpolate(mode='nearest')`` source indices (int32 device tensors)."""
    def one(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        idx = np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64)
        return torch.from_numpy(np.minimum(idx, n_in - 1).astype(np.int32)).to(device)
    return one(in_shape[0], out_shape[0]), one(in_shape[1], out_shape[1])


# ------------------------------------------------------------------------- scratch
def _dev_index(device) -> int:
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else device.index


_scratch = {}


def scratch(kind: str, nbytes: int, device) -> torch.Tensor:
    """Zero-initialised scratch, cached per (device, kind, size, current stream): two streams never share
    accumulators.  Every kernel leaves its scratch zeroed on return, so one allocation serves every call;
    a call that fails may not have, which is why ``_call`` wipes the cache's buffers before it re-raises."""
    key = (_dev_index(device), kind, int(nbytes), torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None:
        buf = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _call(name: str, *args):
    """``_lib.call`` for the entry points that use ``scratch`` buffers: a failed call may leave counters behind."""
    try:
        _lib.call(name, *args)
    except _lib.EvrepError:
        for buf in _scratch.values():
            buf.zero_()
        raise


_workspace = {}


def workspace(kind: str, nbytes: int, device) -> torch.Tensor:
    """Grow-only uninitialised scratch: one buffer per (device, kind, current stream), so that two
    streams never share one and a buffer is only ever released by the stream that uses it."""
    key = (_dev_index(device), kind, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspace.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspace[key] = buf = None          # release before growing
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspace[key] = buf
    return buf


def _workspace_peek(kind: str, device):
    return _workspace.get((_dev_index(device), kind, torch.cuda.current_stream(device).cuda_stream))


# -------------------------------------------------------------------- event buffers
@dataclass
class EventStream:
    """Structure-of-arrays device buffers of one (slice of a) recording."""
    t: torch.Tensor     # uint32 [n]  microseconds
    x: torch.Tensor     # uint16 [n]
    y: torch.Tensor     # uint16 [n]
    p: torch.Tensor     # uint8  [n]
    ordered: Optional[bool] = None     # timestamps non-decreasing (None = not known yet, see is_ordered)

    @property
    def n(self) -> int:
        return int(self.t.shape[0])

    def is_ordered(self) -> bool:
        """Whether the timestamps are non-decreasing (what ``src/io/psee_loader.py`` assumes of a
        file).  Checked on the device once per stream (one 4-byte read back) and remembered; the
        ``*_ordered`` entry points of the library need it."""
        if self.ordered is None:
            flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            _call("evrep_events_order_check", _ptr(self.t), self.n, _ptr(flag), _stream(self.device))
            self.ordered = int(flag.item()) == 0
        return self.ordered

    @property
    def device(self):
        return self.t.device

    @classmethod
    def from_numpy(cls, t, x, y, p, device="cuda") -> "EventStream":
        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
        return cls(up(t, np.uint32), up(x, np.uint16), up(y, np.uint16), up(p, np.uint8))

    @classmethod
    def empty(cls, n: int, device="cuda") -> "EventStream":
        return cls(torch.empty(n, dtype=torch.uint32, device=device), torch.empty(n, dtype=torch.uint16, device=device),
                   torch.empty(n, dtype=torch.uint16, device=device), torch.empty(n, dtype=torch.uint8, device=device))

    def slice(self, lo: int, hi: int) -> "EventStream":
        return EventStream(self.t[lo:hi], self.x[lo:hi], self.y[lo:hi], self.p[lo:hi], True if self.ordered else None)

    def to_aos64(self) -> torch.Tensor:
        """The reference's staging matrix: float64 ``[n,4]`` columns (x, y, t, p)."""
        out = torch.empty((self.n, 4), dtype=torch.float64, device=self.device)
        _call("evrep_soa_to_aos64", _ptr(self.t), _ptr(self.x), _ptr(self.y), _ptr(self.p),
                  self.n, _ptr(out), _stream(self.device))
        return out


def decode_dat(records: torch.Tensor, out: Optional[EventStream] = None) -> EventStream:
    """Decode raw ``.dat`` payload bytes (uint8 CUDA tensor, 8 bytes per event)."""
    _need_cuda(records)
    assert records.dtype == torch.uint8 and records.is_contiguous() and records.numel() % 8 == 0
    n = records.numel() // 8
    ev = out if out is not None else EventStream.empty(n, records.device)
    assert ev.n == n
    _call("evrep_decode_dat", _ptr(records), n, _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p),
              _stream(records.device))
    return ev


def _maps(maps: Optional[CoordMaps]):
    return (_ptr(None), _ptr(None)) if maps is None else (_ptr(maps.xmap), _ptr(maps.ymap))


# ------------------------------------------------------------------------- encoders
def count_accumulate(ev: EventStream, shape, maps=None, counts=None) -> torch.Tensor:
    H, W = shape
    if counts is None:
        counts = scratch("count", 8 * H * W, ev.device).view(torch.int32)
    xm, ym = _maps(maps)
    _call("evrep_count_accumulate", _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n, H, W, xm, ym,
              _ptr(counts), _stream(ev.device))
    return counts


def count_finalize(counts, shape, reset=True, out=None) -> torch.Tensor:
    H, W = shape
    if out is None:
        out = torch.empty((2, H, W), dtype=torch.float32, device=counts.device)
    _call("evrep_count_finalize", _ptr(counts), H, W, _ptr(out), int(bool(reset)), _stream(counts.device))
    return out


def count_image(ev: EventStream, shape, maps=None, out=None) -> torch.Tensor:
    """E1 on a SoA slice: f32 ``[2,H,W]``."""
    _need_cuda(ev.t)
    return count_finalize(count_accumulate(ev, shape, maps), shape, True, out)


def count_images_nested(ev: EventStream, sizes: Sequence[int], shape, maps=None):
    """The driver's last-N windows (``generate_eventcountimage.py:156``) for ascending N,
    accumulating every event once: returns one f32 ``[2,H,W]`` per N."""
    sizes = list(sizes)
    assert sizes == sorted(sizes)
    outs, done = [], 0
    counts = None
    for i, n in enumerate(sizes):
        take = min(n, ev.n)
        if take > done:
            counts = count_accumulate(ev.slice(ev.n - take, ev.n - done), shape, maps, counts)
            done = take
        if counts is None:
            counts = scratch("count", 8 * shape[0] * shape[1], ev.device).view(torch.int32)
        outs.append(count_finalize(counts, shape, reset=(i == len(sizes) - 1)))
    return outs


def count_image_aos64(events: torch.Tensor, shape) -> torch.Tensor:
    _need_cuda(events)
    H, W = shape
    events = events.contiguous()
    counts = scratch("count", 8 * H * W, events.device).view(torch.int32)
    _call("evrep_count_accumulate_aos64", _ptr(events), events.shape[0], events.shape[1], H, W,
              _ptr(counts), _stream(events.device))
    return count_finalize(counts, shape, True)


def _sae_scalars(now):
    now_f32 = np.float32(float(now))
    init = np.float32(now_f32 - np.float32(5000000.0))       # generate_surfaceofactiveevents.py:48
    return float(init), float(now_f32)


def sae(ev: EventStream, shape, lambdas, memory, now, maps=None):
    """A1 on a SoA slice: ``(f32 [2L,H,W], f32 memory [2,H,W])``."""
    _need_cuda(ev.t, memory)
    H, W = shape
    L = len(lambdas)
    lam = (ctypes.c_float * L)(*[float(np.float32(v)) for v in lambdas])
    init, now_f32 = _sae_scalars(now)
    out = torch.empty((2 * L, H, W), dtype=torch.float32, device=ev.device)
    mem_out = torch.empty((2, H, W), dtype=torch.float32, device=ev.device)
    keys = scratch("sae", 8 * H * W, ev.device)
    xm, ym = _maps(maps)
    _call("evrep_sae", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n, H, W, xm, ym, init, now_f32,
              ctypes.cast(lam, ctypes.c_void_p), L, _ptr(memory), _ptr(mem_out), _ptr(keys), _ptr(out),
              _stream(ev.device))
    return out, mem_out


def sae_aos64(events, shape, lambdas, memory, now):
    _need_cuda(events, memory)
    H, W = shape
    L = len(lambdas)
    events = events.contiguous()
    lam = (ctypes.c_float * L)(*[float(np.float32(v)) for v in lambdas])
    init, now_f32 = _sae_scalars(now)
    out = torch.empty((2 * L, H, W), dtype=torch.float32, device=events.device)
    mem_out = torch.empty((2, H, W), dtype=torch.float32, device=events.device)
    keys = scratch("sae", 8 * H * W, events.device)
    _call("evrep_sae_aos64", _ptr(events), events.shape[0], events.shape[1], H, W, init, now_f32,
              ctypes.cast(lam, ctypes.c_void_p), L, _ptr(memory), _ptr(mem_out), _ptr(keys), _ptr(out),
              _stream(events.device))
    return out, mem_out


def event_volume(ev: EventStream, t0: int, tw: int, shape, K: int, maps=None, out=None):
    """V1 on a SoA slice, ``t_norm = (t - t0) / tw``: f32 ``[2K,H,W]``."""
    _need_cuda(ev.t)
    H, W = shape
    if out is None:
        out = torch.empty((2 * K, H, W), dtype=torch.float32, device=ev.device)
    xm, ym = _maps(maps)
    _call("evrep_event_volume", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n, int(t0), int(tw),
              H, W, K, xm, ym, _ptr(out), _stream(ev.device))
    return out


def event_volume_aos64(events, shape, K: int):
    _need_cuda(events)
    H, W = shape
    events = events.contiguous()
    out = torch.empty((2 * K, H, W), dtype=torch.float32, device=events.device)
    _call("evrep_event_volume_aos64", _ptr(events), events.shape[0], events.shape[1], H, W, K, _ptr(out),
              _stream(events.device))
    return out


def taf_fresh_state(shape, K: int, device) -> torch.Tensor:
    return torch.full((shape[0], shape[1], 2, K), TAF_INIT, dtype=torch.float32, device=device)


def _taf_scratch(shape, device):
    nbytes = _lib.load().evrep_taf_bin_scratch_bytes(shape[0], shape[1])
    return scratch("taf_bin", nbytes, device)


def taf_bin(ev: EventStream, t_min: int, t_span: float, shape, K: int, state, maps=None,
            want_out=True, in_place=False):
    """T1 on a SoA slice: ``(f32 [2K,H,W] or None, new state f32 [H,W,2,K])``."""
    _need_cuda(ev.t, state)
    H, W = shape
    state = state.contiguous()
    new_state = state if in_place else torch.empty_like(state)
    out = torch.empty((2 * K, H, W), dtype=torch.float32, device=state.device) if want_out else None
    xm, ym = _maps(maps)
    _call("evrep_taf_bin", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n, int(t_min), float(t_span),
              H, W, K, xm, ym, _ptr(state), _ptr(new_state), _ptr(out), _ptr(_taf_scratch(shape, state.device)),
              _stream(state.device))
    return out, new_state


def taf_bin_aos64(events, shape, K: int, state):
    _need_cuda(events, state)
    H, W = shape
    events = events.contiguous()
    state = state.contiguous()
    new_state = torch.empty_like(state)
    out = torch.empty((2 * K, H, W), dtype=torch.float32, device=state.device)
    _call("evrep_taf_bin_aos64", _ptr(events), events.shape[0], events.shape[1] if events.dim() == 2 else 5,
              H, W, K, _ptr(state), _ptr(new_state), _ptr(out), _ptr(_taf_scratch(shape, state.device)),
              _stream(state.device))
    return out, new_state


TAF_ORDERED_MAX_WINDOW = 128 * 8188      # events per window the one-pass ordered path accepts (kBinMajorMaxParts slices)
_last_status = {}                        # device index -> int32 view of the status block of the last ordered TAF call


def taf_stream(ev: EventStream, windows, abin: int, shape, K: int, state, maps=None,
               emit_state_every_window=False, out=None, tile_events=None, out_u8=None, want_f32=True):
    """T2: run a list of windows ``(ev_begin, ev_end, start_time, n_bins, fresh)`` through the
    whole-stream TAF kernels.  ``state`` (f32 ``[H,W,2,K]``) is updated IN PLACE.  Returns f32
    ``[n_windows, 2K, H, W]`` (``None`` with ``want_f32=False``).  ``out_u8``: optional u8
    ``[n_windows, K, 2, H, W]`` filled with the bytes of the ``bins*`` files (leaky transform +
    slot flip, ``generate_taf.py:226-235``).  ``tile_events``: optional pair of
    ``torch.cuda.Event(enable_timing=True)`` recorded around the tile kernel.

    Three implementations with the same results:
    * default (and ``EVREP_TAF_PATH=bucketed``): the general two-pass bucketing (``evrep_taf_stream``), the
      fastest of the three as measured;
    * ``EVREP_TAF_PATH=ordered``, time-ordered streams (``ev.is_ordered()``) whose windows hold at most
      ``TAF_ORDERED_MAX_WINDOW`` events: one-pass bin-major sort + register-resident tile kernel
      (``evrep_taf_stream_ordered``);
    * ``EVREP_TAF_PATH=sliced``: slice sort + shared-memory tile kernel with lazy ageing
      (``evrep_taf_stream_sliced``), which writes ``out_u8`` itself."""
    _need_cuda(ev.t, state)
    H, W = shape
    nw = len(windows)
    keep, arr = _records([tuple(w[:5]) for w in windows], _lib.TafWindow)
    total_bins = sum(int(w[3]) for w in windows)
    if out is None and want_f32:
        out = torch.empty((nw, 2 * K, H, W), dtype=torch.float32, device=ev.device)
    assert state.is_contiguous() and (out is None or out.is_contiguous())
    xm, ym = _maps(maps)
    sensor = maps.sensor_shape if maps is not None else (H, W)
    lib = _lib.load()
    path = os.environ.get("EVREP_TAF_PATH", "")
    if path == "sliced" and ev.is_ordered():
        need = lib.evrep_taf_stream_sliced_scratch_bytes(ev.n, nw, total_bins, H, W, K)
        if need < 0:
            _lib.check(int(need), "evrep_taf_stream_sliced_scratch_bytes")
        buf = workspace("taf_ordered", need, ev.device)
        if out_u8 is not None:
            assert out_u8.is_contiguous() and out_u8.dtype == torch.uint8 and out_u8.numel() >= nw * 2 * K * H * W
        _call("evrep_taf_stream_sliced", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              arr, nw, int(abin), H, W, K, xm, ym, sensor[0], sensor[1], _ptr(state),
              int(bool(emit_state_every_window)), _ptr(out), 2 * K * H * W, _ptr(out_u8), 2 * K * H * W,
              _ptr(buf), buf.numel(),
              _event(tile_events, 0, ev.device), _event(tile_events, 1, ev.device), _stream(ev.device))
        _last_status[_dev_index(ev.device)] = buf[:16].view(torch.int32)
        return out
    vol = out if out is not None else torch.empty((nw, 2 * K, H, W), dtype=torch.float32, device=ev.device)
    small = all(int(w[1]) - int(w[0]) <= TAF_ORDERED_MAX_WINDOW for w in windows)
    if path == "ordered" and small and ev.is_ordered():
        need = lib.evrep_taf_stream_ordered_scratch_bytes(ev.n, nw, total_bins, H, W)
        if need < 0:
            _lib.check(int(need), "evrep_taf_stream_ordered_scratch_bytes")
        buf = workspace("taf_ordered", need, ev.device)
        _call("evrep_taf_stream_ordered", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              arr, nw, int(abin), H, W, K, xm, ym, sensor[0], sensor[1], _ptr(state),
              int(bool(emit_state_every_window)), _ptr(vol), 2 * K * H * W, _ptr(buf), buf.numel(),
              _event(tile_events, 0, ev.device), _event(tile_events, 1, ev.device), _stream(ev.device))
        at = int(lib.evrep_taf_stream_ordered_status_offset(ev.n, nw, total_bins, H, W))
        _last_status[_dev_index(ev.device)] = buf[at:at + 16].view(torch.int32)
    else:
        need = lib.evrep_taf_stream_scratch_bytes(ev.n, nw, total_bins, H, W)
        if need < 0:
            _lib.check(int(need), "evrep_taf_stream_scratch_bytes")
        buf = workspace("taf_stream", need, ev.device)
        _call("evrep_taf_stream", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              arr, nw, int(abin), H, W, K, xm, ym, sensor[0], sensor[1], _ptr(state),
              int(bool(emit_state_every_window)), _ptr(vol), 2 * K * H * W, _ptr(buf), buf.numel(),
              _event(tile_events, 0, ev.device), _event(tile_events, 1, ev.device), _stream(ev.device))
        _last_status.pop(_dev_index(ev.device), None)
    if out_u8 is not None:
        taf_leaky_u8_batch(vol, K, out=out_u8)
    return out


def order_violations_tensor(device) -> torch.Tensor:
    """Device view (int32 ``[4]``) of the status block of the last ordered ``taf_stream`` call on ``device``: ``[0]`` events
    found outside the bin their position implies, ``[2]`` sort CTAs that gave up waiting for their bin's siblings."""
    view = _last_status.get(_dev_index(device))
    return view if view is not None else torch.zeros(4, dtype=torch.int32, device=device)


def order_violations(device) -> int:
    """Problems the last ``taf_stream`` call on an ordered path reported (0 = the stream really was ordered and the sort
    ran to completion).  Synchronises."""
    st = order_violations_tensor(device).cpu()
    return int(st[0]) + int(st[2])


def _event(pair, i, device):
    if pair is None:
        return ctypes.c_void_p(0)
    e = pair[i]
    if not e.cuda_event:          # the CUDA event is created lazily, on first record
        e.record(torch.cuda.current_stream(device))
    return ctypes.c_void_p(e.cuda_event)


# ------------------------------------------------------------------------ epilogues
def nearest_resize(volume, target_shape, maps=None):
    _need_cuda(volume)
    C, H, W = volume.shape
    Ht, Wt = target_shape
    ys, xs = maps if maps is not None else nearest_maps((H, W), (Ht, Wt), volume.device)
    out = torch.empty((C, Ht, Wt), dtype=torch.float32, device=volume.device)
    _call("evrep_nearest_resize", _ptr(volume.contiguous()), C, H, W, Ht, Wt, _ptr(ys), _ptr(xs), _ptr(out),
              _stream(volume.device))
    return out


def quantize_u8(volume, clamp255=False, out=None):
    _need_cuda(volume)
    volume = volume.contiguous()
    if out is None:
        out = torch.empty(volume.shape, dtype=torch.uint8, device=volume.device)
    _call("evrep_quantize_u8", _ptr(volume), volume.numel(), int(bool(clamp255)), _ptr(out), _stream(volume.device))
    return out


def leaky_transform(ecd):
    _need_cuda(ecd)
    src = ecd.contiguous()
    out = torch.empty_like(src)
    _call("evrep_leaky_transform", _ptr(src), src.numel(), _ptr(out), _stream(src.device))
    return out


def taf_leaky_u8(volume, K: int, target_shape=None, maps=None, out=None):
    """``generate_taf.py:226-235`` fused: f32 ``[2K,H,W]`` -> u8 ``[K,2,Ht,Wt]``, slot 0 newest."""
    _need_cuda(volume)
    C, H, W = volume.shape
    assert C == 2 * K
    Ht, Wt = target_shape if target_shape is not None else (H, W)
    if (Ht, Wt) != (H, W) and maps is None:
        maps = nearest_maps((H, W), (Ht, Wt), volume.device)
    ys, xs = maps if maps is not None else (None, None)
    if out is None:
        out = torch.empty((K, 2, Ht, Wt), dtype=torch.uint8, device=volume.device)
    _call("evrep_taf_leaky_u8", _ptr(volume.contiguous()), K, H, W, Ht, Wt, _ptr(ys), _ptr(xs), _ptr(out),
              _stream(volume.device))
    return out


def taf_leaky_u8_batch(volumes, K: int, target_shape=None, maps=None, out=None):
    """All windows of a ``taf_stream`` result at once: f32 ``[n,2K,H,W]`` -> u8 ``[n,K,2,Ht,Wt]``."""
    _need_cuda(volumes)
    n, C, H, W = volumes.shape
    assert C == 2 * K and volumes.is_contiguous()
    Ht, Wt = target_shape if target_shape is not None else (H, W)
    if (Ht, Wt) != (H, W) and maps is None:
        maps = nearest_maps((H, W), (Ht, Wt), volumes.device)
    ys, xs = maps if maps is not None else (None, None)
    if out is None:
        out = torch.empty((n, K, 2, Ht, Wt), dtype=torch.uint8, device=volumes.device)
    _call("evrep_taf_leaky_u8_batch", _ptr(volumes), C * H * W, n, K, H, W, Ht, Wt, _ptr(ys), _ptr(xs), _ptr(out),
              _stream(volumes.device))
    return out


def event_volume_stream(ev: EventStream, windows, tw: int, shape, K: int, maps=None, out=None):
    """V2: Event Volumes of a list of ordered, non-overlapping windows ``(ev_begin, ev_end, t0)``
    of common length ``tw`` in one call.  Returns f32 ``[n_windows, 2K, H, W]``."""
    _need_cuda(ev.t)
    H, W = shape
    nw = len(windows)
    keep, arr = _records([tuple(w[:3]) for w in windows], _lib.EvWindow)
    if out is None:
        out = torch.empty((nw, 2 * K, H, W), dtype=torch.float32, device=ev.device)
    assert out.is_contiguous()
    need = _lib.load().evrep_event_volume_stream_scratch_bytes(ev.n, nw, H, W)
    if need < 0:
        _lib.check(int(need), "evrep_event_volume_stream_scratch_bytes")
    buf = workspace("taf_stream", need, ev.device)
    xm, ym = _maps(maps)
    _call("evrep_event_volume_stream", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              arr, nw, int(tw), H, W, K, xm, ym,
              maps.sensor_shape[0] if maps is not None else H, maps.sensor_shape[1] if maps is not None else W,
              _ptr(out), 2 * K * H * W, _ptr(buf), buf.numel(), _stream(ev.device))
    return out


EV_SEGMENT_MAX_US = 262000      # a record keeps an 18-bit offset from its segment's start


def plan_ev_spans(windows, time_of, lower_index):
    """Host-side plan of ``event_volume_spans``.  ``windows``: ``(ev_begin, ev_end, t0, tw)`` in any
    order, free to overlap and nest (``generate_eventvolume.py:118-147``).  ``time_of(i)``: timestamp of
    event ``i``; ``lower_index(T, lo, hi)``: first index in ``[lo, hi)`` with ``t >= T`` (both on the
    host, e.g. ``PSEELoader.time_of`` / ``lower_index``).  The stream is cut at every window boundary,
    and wherever a piece would span more than ``EV_SEGMENT_MAX_US``; pieces no window covers are
    left out.  Returns ``(segments [(ev_begin, ev_end, start_time)], spans [(first, last, t0, tw)])``."""
    marks = {}
    for a, b, _t0, _tw in windows:
        if b > a:
            marks[int(a)] = marks.get(int(a), 0) + 1
            marks[int(b)] = marks.get(int(b), 0) - 1
    cuts = sorted(marks)
    segments, first_at, last_at = [], {}, {}
    cover = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        cover += marks[lo]
        if cover <= 0:
            continue
        first_at[lo] = len(segments)
        while True:
            start = int(time_of(lo))
            if int(time_of(hi - 1)) - start <= EV_SEGMENT_MAX_US:
                break
            cut = int(lower_index(start + EV_SEGMENT_MAX_US, lo, hi))      # > lo: t[lo] = start < start + max
            segments.append((lo, cut, start))
            lo = cut
        segments.append((lo, hi, start))
        last_at[hi] = len(segments) - 1
    spans = []
    for a, b, t0, tw in windows:
        if b > a:
            spans.append((first_at[int(a)], last_at[int(b)], int(t0), int(tw)))
        else:
            spans.append((0, -1, int(t0), int(tw)))
    return segments, spans


def event_volume_spans(ev: EventStream, segments, spans, shape, K: int, maps=None, out=None, out_u8=None,
                       want_f32=True, tile_events=None):
    """V2 for overlapping / nested windows on a time-ordered stream (``evrep_event_volume_spans``): the
    events are sorted once by (segment, sensor tile), every span splats the records of its segments with
    its own ``(t - t0) / tw``.  ``segments`` / ``spans`` as returned by ``plan_ev_spans``.  Returns f32
    ``[n_spans, 2K, H, W]`` (``None`` with ``want_f32=False``); ``out_u8``: optional u8 tensor of the
    same shape that receives the file bytes (clamp at 255, truncation) when no resize follows."""
    _need_cuda(ev.t)
    H, W = shape
    ns, nsp = len(segments), len(spans)
    keep_seg, seg_arr = _records(segments, _lib.EvSegment)
    keep_span, span_arr = _records(spans, _lib.EvSpan)
    if out is None and want_f32:
        out = torch.empty((nsp, 2 * K, H, W), dtype=torch.float32, device=ev.device)
    need = _lib.load().evrep_event_volume_spans_scratch_bytes(ev.n, ns, nsp, H, W, K)
    if need < 0:
        _lib.check(int(need), "evrep_event_volume_spans_scratch_bytes")
    buf = workspace("ev_spans", need, ev.device)
    xm, ym = _maps(maps)
    sensor = maps.sensor_shape if maps is not None else (H, W)
    _call("evrep_event_volume_spans", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              seg_arr, ns, span_arr, nsp, H, W, K, xm, ym,
              sensor[0], sensor[1], _ptr(out), 2 * K * H * W, _ptr(out_u8), 2 * K * H * W, _ptr(buf), buf.numel(),
              _event(tile_events, 0, ev.device), _event(tile_events, 1, ev.device), _stream(ev.device))
    return out


def event_volume_u8_batch(volumes, target_shape=None, maps=None, out=None):
    """f32 ``[n,C,H,W]`` -> u8 ``[n,C,Ht,Wt]``: optional nearest resize (gen1 policy), clamp at 255,
    truncation (``generate_eventvolume.py:150-160`` for a batch of windows)."""
    _need_cuda(volumes)
    n, C, H, W = volumes.shape
    assert volumes.is_contiguous()
    Ht, Wt = target_shape if target_shape is not None else (H, W)
    if (Ht, Wt) != (H, W) and maps is None:
        maps = nearest_maps((H, W), (Ht, Wt), volumes.device)
    ys, xs = maps if (maps is not None and (Ht, Wt) != (H, W)) else (None, None)
    if out is None:
        out = torch.empty((n, C, Ht, Wt), dtype=torch.uint8, device=volumes.device)
    _call("evrep_event_volume_u8_batch", _ptr(volumes), C * H * W, n, C, H, W, Ht, Wt, _ptr(ys), _ptr(xs), _ptr(out),
              _stream(volumes.device))
    return out


def timesurface(ev: EventStream, shape):
    """Time-surface pair of ``generate_opticalflow.py:72-92`` on a SoA slice: ``(f64 [H,W] last
    timestamp older than newest - 50000, f64 [H,W] last timestamp)``, shifted, scaled and clamped
    like the reference."""
    _need_cuda(ev.t)
    H, W = shape
    out_old = torch.empty((H, W), dtype=torch.float64, device=ev.device)
    out_all = torch.empty((H, W), dtype=torch.float64, device=ev.device)
    buf = scratch("timesurface", _lib.load().evrep_timesurface_scratch_bytes(H, W), ev.device)
    _call("evrep_timesurface", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), ev.n, H, W, _ptr(buf), _ptr(out_old), _ptr(out_all),
              _stream(ev.device))
    return out_old, out_all


def plan_count_segments(windows):
    """Host-side plan of ``count_stream``: cut the event axis at every boundary of the non-empty
    windows.  Returns ``(segments, order, emits)``: consecutive ``(ev_begin, ev_end)`` segments,
    the indices of the non-empty windows sorted by their last segment, and for each of those (in
    that order) the inclusive run ``(first_segment, last_segment)`` it covers."""
    w = np.asarray(windows, dtype=np.int64).reshape(-1, 2)
    live = np.nonzero(w[:, 1] > w[:, 0])[0]
    if live.size == 0:
        return [], [], []
    bounds = np.unique(w[live])
    first = np.searchsorted(bounds, w[live, 0])
    last = np.searchsorted(bounds, w[live, 1]) - 1
    by_last = np.argsort(last, kind="stable")
    segments = list(zip(bounds[:-1].tolist(), bounds[1:].tolist()))
    return segments, live[by_last].tolist(), list(zip(first[by_last].tolist(), last[by_last].tolist()))


def count_stream(ev: EventStream, windows, shape, maps=None):
    """E1 + E2 for many windows in one call.  ``windows``: ``(ev_begin, ev_end)`` event ranges that
    may nest and overlap (the driver's last-N windows of consecutive labels).  Returns u8
    ``[n_windows, 2, H, W]`` event counts saturated at 255 (the image value depends on
    ``min(count, 20)`` only); feed it to ``count_lut_u8_batch``."""
    _need_cuda(ev.x)
    H, W = shape
    nw = len(windows)
    frames = torch.zeros((nw, 2, H, W), dtype=torch.uint8, device=ev.device)
    segments, order, runs = plan_count_segments(windows)
    if not order:
        return frames
    n_seg = len(segments)
    keep_seg, seg = _records(segments, _lib.CountSegment)
    keep_emits, emits = _records(runs, _lib.CountEmit)
    out = frames if order == list(range(nw)) else torch.empty((len(order), 2, H, W), dtype=torch.uint8, device=ev.device)
    need = _lib.load().evrep_count_stream_scratch_bytes(ev.n, n_seg, len(order), H, W)
    if need < 0:
        _lib.check(int(need), "evrep_count_stream_scratch_bytes")
    buf = workspace("taf_stream", need, ev.device)
    xm, ym = _maps(maps)
    _call("evrep_count_stream", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              seg, n_seg, emits, len(order), H, W, xm, ym,
              maps.sensor_shape[0] if maps is not None else H, maps.sensor_shape[1] if maps is not None else W,
              _ptr(out), 2 * H * W, _ptr(buf), buf.numel(), _stream(ev.device))
    if out is not frames:
        frames[torch.as_tensor(order, device=ev.device)] = out
    return frames


def count_lut_u8_batch(frames, target_shape=None, resize_maps=None, out=None):
    """Value LUT, nearest resize and uint8 truncation of the count-image driver for all frames of a
    ``count_stream`` result: u8 ``[n, 2, H, W]`` counts -> u8 ``[n, 2, Ht, Wt]`` images."""
    _need_cuda(frames)
    n, two, H, W = frames.shape
    assert two == 2 and frames.is_contiguous()
    Ht, Wt = target_shape if target_shape is not None else (H, W)
    if (Ht, Wt) != (H, W) and resize_maps is None:
        resize_maps = nearest_maps((H, W), (Ht, Wt), frames.device)
    ys, xs = resize_maps if (Ht, Wt) != (H, W) else (None, None)
    if out is None:
        out = torch.empty((n, 2, Ht, Wt), dtype=torch.uint8, device=frames.device)
    _call("evrep_count_lut_u8_batch", _ptr(frames), 2 * H * W, n, H, W, Ht, Wt, _ptr(ys), _ptr(xs), _ptr(out),
              _stream(frames.device))
    return out


def sae_stream(ev: EventStream, windows, shape, memory=None, maps=None, out=None):
    """A1 + A2 for every label of a recording in one call.  ``windows``: ordered, non-overlapping
    ``(ev_begin, ev_end, now, t_first, t_last)`` -- the events the driver hands to the encoder for
    label ``now`` and the timestamps of the first / last of them.  Returns ``(latest f32
    [n_windows, 2, H, W], memory f32 [2, H, W])``: ``latest[w]`` is the reference's ``t_img`` after
    the merge with ``memory`` (generate_surfaceofactiveevents.py:52), ``memory`` the state after
    the last window (a new tensor, the argument is not modified)."""
    _need_cuda(ev.t, memory)
    H, W = shape
    nw = len(windows)
    keep, arr = _records([tuple(w[:5]) for w in windows], _lib.SaeWindow)
    if out is None:
        out = torch.empty((nw, 2, H, W), dtype=torch.float32, device=ev.device)
    assert out.is_contiguous()
    state = memory.clone() if memory is not None else torch.empty((2, H, W), dtype=torch.float32, device=ev.device)
    need = _lib.load().evrep_sae_stream_scratch_bytes(ev.n, arr, nw, H, W)
    if need < 0:
        _lib.check(int(need), "evrep_sae_stream_scratch_bytes")
    buf = workspace("taf_stream", need, ev.device)
    xm, ym = _maps(maps)
    _call("evrep_sae_stream", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n,
              arr, nw, H, W, xm, ym,
              maps.sensor_shape[0] if maps is not None else H, maps.sensor_shape[1] if maps is not None else W,
              _ptr(state), 1 if memory is not None else 0, _ptr(out), 2 * H * W, _ptr(buf), buf.numel(), _stream(ev.device))
    return out, state


def sae_decay_u8_batch(latest, nows, lambdas, target_shape=None, resize_maps=None, out=None):
    """The decays, nearest resize and uint8 truncation of the SAE driver for all windows of a
    ``sae_stream`` result: f32 ``[n, 2, H, W]`` -> u8 ``[n, L, 2, Ht, Wt]``."""
    _need_cuda(latest)
    n, two, H, W = latest.shape
    assert two == 2 and latest.is_contiguous()
    Ht, Wt = target_shape if target_shape is not None else (H, W)
    if (Ht, Wt) != (H, W) and resize_maps is None:
        resize_maps = nearest_maps((H, W), (Ht, Wt), latest.device)
    ys, xs = resize_maps if (Ht, Wt) != (H, W) else (None, None)
    L = len(lambdas)
    lam = (ctypes.c_float * L)(*[float(np.float32(v)) for v in lambdas])
    now_f32 = torch.from_numpy(np.asarray([float(v) for v in nows], dtype=np.float64).astype(np.float32)).to(latest.device)
    if out is None:
        out = torch.empty((n, L, 2, Ht, Wt), dtype=torch.uint8, device=latest.device)
    _call("evrep_sae_decay_u8_batch", _ptr(latest), 2 * H * W, _ptr(now_f32), n, H, W, Ht, Wt, _ptr(ys), _ptr(xs),
              ctypes.cast(lam, ctypes.c_void_p), L, _ptr(out), _stream(latest.device))
    return out


def count_images_u8(ev: EventStream, sizes: Sequence[int], shape, target_shape, maps=None, resize_maps=None, out=None):
    """Driver form of the count image: the nested last-N windows of one label as uint8
    ``[len(sizes), 2, Ht, Wt]`` in one library call (LUT + nearest resize + truncation fused)."""
    _need_cuda(ev.x)
    H, W = shape
    Ht, Wt = target_shape
    sizes = [int(v) for v in sizes]
    arr = (ctypes.c_int64 * len(sizes))(*sizes)
    if (Ht, Wt) != (H, W) and resize_maps is None:
        resize_maps = nearest_maps((H, W), (Ht, Wt), ev.device)
    ys, xs = resize_maps if (Ht, Wt) != (H, W) else (None, None)
    if out is None:
        out = torch.empty((len(sizes), 2, Ht, Wt), dtype=torch.uint8, device=ev.device)
    counts = scratch("count", 8 * H * W, ev.device)
    xm, ym = _maps(maps)
    _call("evrep_count_images_u8", _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n, ctypes.cast(arr, ctypes.c_void_p),
              len(sizes), H, W, xm, ym, Ht, Wt, _ptr(ys), _ptr(xs), _ptr(counts), _ptr(out), _stream(ev.device))
    return out


def sae_u8(ev: EventStream, shape, target_shape, lambdas, memory, now, maps=None, resize_maps=None, out=None):
    """Driver form of the SAE: ``(uint8 [L,2,Ht,Wt], new memory f32 [2,H,W])`` in one library call."""
    _need_cuda(ev.t, memory)
    H, W = shape
    Ht, Wt = target_shape
    L = len(lambdas)
    lam = (ctypes.c_float * L)(*[float(np.float32(v)) for v in lambdas])
    init, now_f32 = _sae_scalars(now)
    if (Ht, Wt) != (H, W) and resize_maps is None:
        resize_maps = nearest_maps((H, W), (Ht, Wt), ev.device)
    ys, xs = resize_maps if (Ht, Wt) != (H, W) else (None, None)
    if out is None:
        out = torch.empty((L, 2, Ht, Wt), dtype=torch.uint8, device=ev.device)
    mem_out = torch.empty((2, H, W), dtype=torch.float32, device=ev.device)
    keys = scratch("sae", 8 * H * W, ev.device)
    xm, ym = _maps(maps)
    _call("evrep_sae_u8", _ptr(ev.t), _ptr(ev.x), _ptr(ev.y), _ptr(ev.p), ev.n, H, W, xm, ym, Ht, Wt, _ptr(ys), _ptr(xs),
              init, now_f32, ctypes.cast(lam, ctypes.c_void_p), L, _ptr(memory), _ptr(mem_out), _ptr(keys), _ptr(out),
              _stream(ev.device))
    return out, mem_out
