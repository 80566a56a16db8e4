"""CPU restatement of the reference's per-window encoders (torch CPU / numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity: pinned against the
unmodified reference by ``oracle/make_golden.py`` -> ``tests/golden/`` and, where
``/root/reference`` is mounted, by ``tests/test_oracle_vs_reference.py``.

Every function keeps the reference's precision sequence (float64 inputs -> the same
casts -> float32 arithmetic in the same order), so integer-valued results are
bit-exact and float results are bit-identical on CPU.  Citations are relative to
``/root/reference``.

Conventions: ``events`` is a float64 ``[N,4]`` tensor ``(x, y, t, p)`` (TAF: ``[N,5]``
with a bin-id column), ``shape = (H, W)``.
"""
from __future__ import annotations

import numpy as np
import torch

F32 = torch.float32


def _cols(events: torch.Tensor):
    events = torch.as_tensor(events)
    return [events[:, i] for i in range(events.shape[1])]


# --------------------------------------------------------------------------- E1
def count_image(events, shape):
    """Event Count Image.  Follows ``generate_eventcountimage.py:19-41``:
    every event adds float32(0.05) to cell ``2x + 2W y + p`` (:32), cells above 1 are
    clamped to 1 (:34), layout ``[H,W,2] -> [2,H,W]`` (:36), scale 255 (:41)."""
    H, W = shape
    x, y, _, p = _cols(events)
    cell = 2 * x.long() + 2 * W * y.long() + p.long()
    acc = torch.zeros(H * W * 2, dtype=F32)
    acc.index_add_(0, cell, torch.full((cell.numel(),), 0.05, dtype=F32))
    acc = torch.where(acc > 1, torch.ones_like(acc), acc)
    return acc.view(H, W, 2).permute(2, 0, 1).contiguous() * 255


def count_image_lut(max_count: int = 32) -> np.ndarray:
    """Value of a count-image cell as a function of its event count:
    the float32 running sum of 0.05, clamped, times 255."""
    lut = np.zeros(max_count + 1, dtype=np.float32)
    s = np.float32(0.0)
    for n in range(1, max_count + 1):
        s = np.float32(s + np.float32(0.05))
        lut[n] = np.float32((np.float32(1.0) if s > 1 else s) * np.float32(255.0))
    return lut


# --------------------------------------------------------------------------- A1
def sae_surfaces(events, shape, lambdas, memory, now):
    """Surface of Active Events.  Follows ``generate_surfaceofactiveevents.py:71-80``
    (bounds filter :72, casts :76) and ``:44-69`` (initial surface ``f32(now) - 5e6``
    :48, last-writer scatter :49 -- sequential on CPU, i.e. the max timestamp for
    time-sorted input -- max-merge with ``memory`` :51-52, state = absolute f32
    timestamps :54, ``exp(lambda * (t - now)) * 255`` per lambda :55-63)."""
    H, W = shape
    events = torch.as_tensor(events)
    keep = (events[:, 0] < W) & (events[:, 1] < H)
    x, y, t, p = _cols(events[keep])
    latest = torch.zeros((2, H, W), dtype=F32) + now - 5000000
    # numpy fancy assignment writes in index order (last duplicate wins) -- the
    # sequential semantics of the reference's CPU index_put_.
    latest.numpy()[p.long().numpy(), y.long().numpy(), x.long().numpy()] = t.float().numpy()
    if memory is not None:
        latest = torch.where(latest > memory, latest, memory)
    state = latest
    rel = latest - now
    planes = torch.stack([torch.exp(lam * rel) for lam in lambdas], 0)
    return planes.view(len(lambdas) * 2, H, W) * 255, state


# --------------------------------------------------------------------------- V1
def event_volume(events, shape, volume_bins=5):
    """Event Volume (temporal bilinear splat).  Follows
    ``generate_eventvolume.py:15-42``: ``t* = K * f32(t_norm)`` (:23), bin centres
    1..K (:27), weight ``1 - |c - t*|`` times ``[p, 1-p]`` kept when >= 0 (:28-29),
    row-wise accumulate at pixel ``x + W y`` (:31-32), ``[H,W,2K] -> [2K,H,W]`` (:35),
    ``/ 5 * 255`` whatever K is (:37).  Channel of an event = ``2k + (1 - p)``."""
    H, W = shape
    K = int(volume_bins)
    x, y, t, p = _cols(events)
    x, y, p = x.long(), y.long(), p.long()
    t_star = (K * t.float())[:, None, None]
    centres = (torch.arange(K)[:, None].expand(K, 2) + 1)[None]
    pol = torch.stack([p, 1 - p], dim=1)[:, None, :]
    w = (1 - torch.abs(centres - t_star)) * pol
    w = torch.where(w >= 0, w, torch.zeros_like(w)).reshape(-1, 2 * K)
    acc = torch.zeros((H * W, 2 * K), dtype=F32)
    acc.index_add_(0, x + W * y, w)
    vol = acc.view(H, W, 2 * K).permute(2, 0, 1).contiguous()
    return vol / 5 * 255


# --------------------------------------------------------------------------- T1
TAF_INIT = -6000.0  # generate_taf.py:207-209


def taf_fresh_state(shape, volume_bins=8):
    H, W = shape
    return torch.zeros((H, W, 2, volume_bins), dtype=F32) + TAF_INIT


def taf_bin_update(events, shape, state, volume_bins=8):
    """One 10 ms TAF step.  Follows ``generate_taf.py:60-67`` (casts) and ``:19-58``:
    per (y, x, p) count and sum of ``f32(t_norm) - 1`` (:23-26), mean ``s / (n + 1e-8)``
    (:27), ``inactive = (n == 0)`` (:35).  No active pixel anywhere -> the state object
    is returned untouched, no ageing (:40-41).  Otherwise every slot ages by 1 and
    active pixels additionally shift their FIFO and push the mean (:43-49).
    Output ``[2K,H,W]``, channel ``2k + p`` (:55).  Returns ``(out, new_state)``."""
    H, W = shape
    K = int(volume_bins)
    cols = _cols(events)
    x, y, t, p = cols[0].long(), cols[1].long(), cols[2].float(), cols[3].long()
    cell = p + 2 * x + 2 * W * y
    n = torch.zeros(H * W * 2, dtype=F32)
    n.index_add_(0, cell, torch.ones(cell.numel(), dtype=F32))
    s = torch.zeros(H * W * 2, dtype=F32)
    s.index_add_(0, cell, t - 1)
    mean = (s / (n + 1e-8)).view(H, W, 2)
    inactive = (n == 0).view(H, W, 2)
    if bool(inactive.all()):
        new_state = state
    else:
        aged = state - 1
        pushed = torch.cat([aged[..., 1:], mean[..., None]], dim=3)
        new_state = torch.where(inactive[..., None], aged, pushed)
    out = new_state.permute(3, 2, 0, 1).contiguous().view(K * 2, H, W)
    return out, new_state


# --------------------------------------------------------------------------- T3
def leaky_transform(ecd):
    """``255 * max(0, 1 - log1p(-v) / 8.7)`` -- ``generate_taf.py:69-76``."""
    v = torch.log1p(-ecd)
    v = 1 - v / 8.7
    v = torch.where(v < 0, torch.zeros_like(v), v)
    return v * 255


# --------------------------------------------------------------------------- R1
def nearest_resize(volume, target_shape):
    """``F.interpolate(volume[None], size=target, mode='nearest')[0]`` as called at
    ``generate_taf.py:222`` and twins (legacy float32 index rule)."""
    return torch.nn.functional.interpolate(volume[None], size=tuple(target_shape), mode="nearest")[0]


def nearest_index_map(n_in: int, n_out: int) -> np.ndarray:
    """Source index of every destination index under the legacy nearest rule:
    ``min(floor(dst * float32(in/out)), in - 1)``."""
    scale = np.float32(n_in) / np.float32(n_out)
    idx = np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, n_in - 1)


def downscale_coordinate_map(n_in: int, ratio: float) -> np.ndarray:
    """gen4 policy: ``coord * ratio`` in float64, truncated by ``.long()``
    (``generate_taf.py:103-104,216-218``)."""
    return (np.arange(n_in, dtype=np.float64) * ratio).astype(np.int64)


# ------------------------------------------------------------------------ S1-S6
def sparse_agile_event_volume(events, B, shape, iter, past_volume=None, events_window=50000,
                              volume_bins=5, infer_time=10000):
    """``data/sparse_ops.py:4-35``.  Full mode (no ``past_volume``): centres 0..K-1,
    ``t* = K t / window``; incremental mode: two fresh bins from
    ``t* = (t - iter + infer_time) / window * K``, the oldest bin of ``past_volume``
    is dropped and its newest bin receives the first fresh bin IN PLACE (:30-31)."""
    H, W = shape
    b, x, y, t, p = [c for c in _cols(events)]
    b, x, y, p = b.long(), x.long(), y.long(), p.long()
    if past_volume is None:
        t_star = (volume_bins * t.float() / events_window)[:, None, None]
        C = volume_bins
    else:
        t_star = ((t.float() - iter + infer_time) / events_window * volume_bins)[:, None, None]
        C = 2
    centres = torch.arange(C)[:, None].expand(C, 2)[None]
    w = (1 - torch.abs(centres - t_star)) * torch.stack([p, 1 - p], dim=1)[:, None, :]
    w = torch.where(w >= 0, w, torch.zeros_like(w)).reshape(-1, 2 * C)
    pix = H * W * b + x + W * y
    if past_volume is None:
        img = torch.zeros((B * H * W, 2 * volume_bins), dtype=F32)
        img.index_add_(0, pix, w)
        img = img.view(B * H * W, volume_bins, 2, 1)
    else:
        fresh = torch.zeros((B * H * W, 4), dtype=F32)
        fresh.index_add_(0, pix, w)
        fresh = fresh.view(B * H * W, 2, 2, 1)
        kept = past_volume[:, 1:]
        kept[:, -1] = kept[:, -1] + fresh[:, 0]
        img = torch.cat([kept, fresh[:, 1:]], dim=1)
    viewed = img.view(B, H, W, img.shape[1] * 2, 1).permute(0, 3, 1, 2, 4).contiguous()
    return viewed, img


def sparse_event_volume(events, B, shape, iter, memory=None, events_window=50000,
                        volume_bins=5, infer_time=10000):
    """``data/sparse_ops.py:37-69``: raw-event memory (:40-42), ``t* = (K-1) t / window``."""
    H, W = shape
    events = torch.as_tensor(events)
    if memory is not None:
        events = torch.cat([memory, events])
    memory = events[events[:, 3] >= iter - events_window + infer_time]
    b, x, y, t, p = _cols(events)
    b, x, y, p = b.long(), x.long(), y.long(), p.long()
    t_star = ((volume_bins - 1) * t.float() / events_window)[:, None, None]
    C = volume_bins
    centres = torch.arange(C)[:, None].expand(C, 2)[None]
    w = (1 - torch.abs(centres - t_star)) * torch.stack([p, 1 - p], dim=1)[:, None, :]
    w = torch.where(w >= 0, w, torch.zeros_like(w)).reshape(-1, 2 * C)
    img = torch.zeros((B * H * W, 2 * C), dtype=F32)
    img.index_add_(0, H * W * b + x + W * y, w)
    img = img.view(B * H * W, C, 2, 1)
    viewed = img.view(B, H, W, 2 * C, 1).permute(0, 3, 1, 2, 4).contiguous()
    return viewed, memory


def sparse_taf(events, B, shape, iter, past_volume=None, events_window=50000,
               volume_bins=5, infer_time=10000):
    """``data/sparse_ops.py:72-85``: scatter-add ``feature`` at (b, c, y, x, p); then
    polarity plane 1 becomes ``-1e8`` where zero and ``+1`` elsewhere (:84)."""
    b, x, y, t, c, p, f = _cols(events)
    b, x, y, p, c = b.long(), x.long(), y.long(), p.long(), c.long()
    H, W = shape
    C = volume_bins * 2
    fmap = torch.zeros(B * C * H * W * 2, dtype=F32)
    fmap.index_add_(0, b * C * H * W * 2 + c * H * W * 2 + y * W * 2 + x * 2 + p, f.float())
    vol = fmap.view(B, C, H, W, 2).contiguous()
    plane = vol[..., 1]
    vol[..., 1] = torch.where(plane == 0, torch.full_like(plane, -1e8), plane + 1)
    return vol, None


def sparse_event_frame(events, B, shape, iter, past_volume=None, events_window=50000,
                       volume_bins=5, infer_time=10000):
    """``data/sparse_ops.py:88-107``: polarity-agnostic occupancy, 255 where any event."""
    H, W = shape
    b, x, y, t, p = _cols(events)
    b, x, y = b.long(), x.long(), y.long()
    img = torch.zeros(B * H * W, dtype=F32)
    img.index_add_(0, H * W * b + x + W * y, torch.ones(b.numel(), dtype=F32))
    img = torch.where(img > 0, torch.full_like(img, 255.0), img)
    img = torch.cat([img, img])
    return img.view(2, B, H, W, 1).permute(1, 0, 2, 3, 4).contiguous(), None


def sparse_to_dense(locations, features, shape):
    """``data/sparse_ops.py:109-121``."""
    B, H, W = shape
    C = features.shape[-1]
    b, y, x = [locations[:, i].long() for i in range(3)]
    fmap = torch.zeros((B * H * W, C), dtype=F32)
    fmap.index_add_(0, H * W * b + W * y + x, features)
    return fmap.view(B, H, W, C)


def dense_to_sparse(dense):
    """``data/sparse_ops.py:123-135``: rows with a non-zero |.|-sum; locations are
    ``(y, x, b)`` (batch index moved last, :131)."""
    nz = torch.nonzero(torch.abs(dense).sum(dim=-1))
    locations = torch.cat((nz[:, 1:], nz[:, 0, None]), dim=-1)
    feats = dense[nz[:, 0], nz[:, 1], nz[:, 2]]
    return locations, feats


# --------------------------------------------------------------------------- N1
def event_queue_tensor(events, queue_length, B, H, W, start_times, event_window_abin):
    """Scalar restatement of ``data/event_representation_tool/src/event_queue_tensor.cpp:10-118``
    AS IT BEHAVES: the per-cell deques are only pushed while non-empty (:53-59), so during
    the event loop every event takes the ``else`` at :69 and adds
    ``1 - (start[b] + abin (z + 1) - t) / abin`` (float32, sequential) to its cell; cells
    with a positive total are then pushed once (:79-91) and land in queue slot Q-1 of
    plane 0, with the bin plane (plane 1) holding the initial -1 there; everything else is
    0 / -1.  Returned as float64 ``[2, Q, 2, B, H, W]`` (the binding declares
    ``array_t<double>``, :10)."""
    ev = np.ascontiguousarray(np.asarray(events, dtype=np.float32))
    start = np.asarray(start_times, dtype=np.int32)
    Q = int(queue_length)
    cells = 2 * B * H * W
    total = np.zeros(cells, dtype=np.float32)
    abin = int(event_window_abin)
    for row in ev:
        b, w, h, p, z = int(row[0]), int(row[1]), int(row[2]), int(row[4]), int(row[5])
        t = np.float32(row[3])
        cell = B * H * W * p + H * W * b + W * h + w
        # int + int*int -> int, then (int - float) -> float32, / int -> float32
        edge = np.float32(int(start[b]) + abin * (z + 1))
        total[cell] = np.float32(total[cell] + np.float32(1 - np.float32(np.float32(edge - t) / np.float32(abin))))
    out = np.zeros((2, Q, cells), dtype=np.float32)
    out[1] = -1.0
    pos = total > 0
    out[0, Q - 1, pos] = total[pos]
    # plane 1 keeps -1 in slot Q-1 for pushed cells too (ecd_now is never updated)
    return out.reshape(2, Q, 2, B, H, W).astype(np.float64)


# --------------------------------------------------------------------------- 8f rank 4
def timesurface_pair(events, shape):
    """Time-surface pair of ``generate_opticalflow.py:72-92`` (``generate_timesurface`` with
    zero-initialised ``volume1`` / ``volume2``): float64 ``[N,4]`` (x, y, t, p) in array order ->
    two float64 ``[H,W]`` surfaces.  ``volume2`` holds the last timestamp per pixel, ``volume1`` the
    last one older than ``end - 50000``; both are shifted to the window start, scaled by
    ``255 / (end - 50000 - start)`` and clamped below at 0.  Polarity is ignored."""
    H, W = shape
    ev = np.asarray(events, dtype=np.float64)
    v1, v2 = np.zeros((H, W)), np.zeros((H, W))
    if len(ev) > 0:
        end, start = ev[:, 2].max(), ev[:, 2].min()
        xs, ys, ts = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 2]
        old = ts < end - 50000
        v1[ys[old], xs[old]] = ts[old]           # fancy assignment writes in array order: last event wins
        v2[ys, xs] = ts
        v1 = v1 - start
        v2 = v2 - start - 50000
        v1 = v1 / (end - 50000 - start) * 255
        v2 = v2 / (end - 50000 - start) * 255
        v1 = np.where(v1 < 0, 0, v1)
        v2 = np.where(v2 < 0, 0, v2)
    return v1, v2
