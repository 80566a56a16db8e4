"""Event Volume for nested / overlapping windows on the slice sort (`evrep_event_volume_spans`): the
reference driver's 250 / 500 / 1000 ms windows per label against the oracle encoder and the per-window
CUDA path, long windows (beyond the 18-bit record offset), the uint8 outputs and the host-side plan.
Float tolerance 1e-5 rel / 1e-6 abs."""
import bisect

import numpy as np
import pytest
import torch

from frlw_evd_b200 import generate_eventvolume as gev
from frlw_evd_b200 import ops, synth
from oracle import encoders as oe

from helpers import staged

pytestmark = pytest.mark.gpu
DEV = "cuda"


def close(a, b):
    return np.allclose(a.cpu().numpy(), b.cpu().numpy() if torch.is_tensor(b) else b, rtol=1e-5, atol=1e-6)


def driver_windows(t, ends, sizes=(250000, 500000, 1000000)):
    """(first, hi, t0, tw) like generate_eventvolume.py:128-147 on an in-memory stream."""
    out = []
    for end in ends:
        hi = bisect.bisect_left(t, end)
        lo = bisect.bisect_left(t, max(end - max(sizes), 0))
        for tw in sizes:
            out.append((bisect.bisect_right(t, end - tw, lo, hi), hi, end - tw, tw))
    return out


def oracle_volume(t, x, y, p, win, grid, K, scale=None):
    lo, hi, t0, tw = win
    e = staged(t, x, y, p, lo, hi)
    e[:, 2] = (e[:, 2] - t0) / tw
    if scale is not None:
        e[:, 0] *= scale[0]
        e[:, 1] *= scale[1]
    return oe.event_volume(e, grid, K)


def host_index(t):
    return (lambda i: int(t[i])), (lambda T, lo, hi: bisect.bisect_left(t, T, lo, hi))


@pytest.mark.parametrize("K", [5, 8])
def test_nested_overlapping_windows_gen1(K):
    """Six labels 50-400 ms apart, three nested windows each (some reach before the first event)."""
    H, W = 240, 304
    t, x, y, p = synth.make_stream(H, W, 1_600_000, 4e5, 7)
    windows = driver_windows(t, [300_000, 700_000, 750_000, 1_100_000, 1_500_000, 1_599_000])
    segments, spans = ops.plan_ev_spans(windows, *host_index(t))
    assert all(hi - lo > 0 and int(t[hi - 1]) - s0 <= ops.EV_SEGMENT_MAX_US and int(t[lo]) >= s0 for lo, hi, s0 in segments)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    got = ops.event_volume_spans(ev, segments, spans, (H, W), K)
    for i, w in enumerate(windows):
        assert close(got[i], oracle_volume(t, x, y, p, w, (H, W), K)), (i, w)
    again = ops.event_volume_spans(ev, segments, spans, (H, W), K)
    assert torch.equal(got, again)                      # fixed-point sums: order independent


def test_gen4_policy_and_uint8_bytes():
    """Down-scaled grid; the kernel's own uint8 output == clamp + truncation of its float output."""
    t, x, y, p = synth.make_stream(720, 1280, 700_000, 4e6, 8)
    windows = driver_windows(t, [400_000, 450_000, 690_000], sizes=(100_000, 200_000, 400_000))
    segments, spans = ops.plan_ev_spans(windows, *host_index(t))
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    K, grid = 5, (512, 640)
    u8 = torch.zeros((len(windows), 2 * K, *grid), dtype=torch.uint8, device=DEV)
    got = ops.event_volume_spans(ev, segments, spans, grid, K, maps, out_u8=u8)
    for i in (0, 4, 8):
        assert close(got[i], oracle_volume(t, x, y, p, windows[i], grid, K, scale=(0.5, 512 / 720))), i
    for i, (lo, hi, t0, tw) in enumerate(windows):
        assert close(got[i], ops.event_volume(ev.slice(lo, hi), t0, tw, grid, K, maps)), i
    assert torch.equal(u8, ops.quantize_u8(got, clamp255=True))
    assert torch.equal(u8, ops.event_volume_u8_batch(got))
    only = torch.zeros_like(u8)
    assert ops.event_volume_spans(ev, segments, spans, grid, K, maps, out_u8=only, want_f32=False) is None
    assert torch.equal(only, u8)


def test_empty_windows_hot_pixel_and_resize_epilogue():
    H, W, K = 240, 304, 5
    rng = np.random.default_rng(3)
    n = 80_000
    t = np.sort(rng.integers(100_000, 900_000, n)).astype(np.uint32)
    x = rng.integers(0, W, n).astype(np.uint16); y = rng.integers(0, H, n).astype(np.uint16); p = rng.integers(0, 2, n).astype(np.uint8)
    hot = rng.random(n) < 0.4
    x[hot], y[hot] = 11, 200                            # the fixed-point word of this cell wraps many times
    windows = [(0, 0, 0, 50_000),                       # before the first event: empty
               (0, n, 100_000, 800_000),
               (bisect.bisect_right(t, 500_000), n, 500_000, 400_000),
               (n, n, 900_000, 250_000)]                # after the last event: empty
    segments, spans = ops.plan_ev_spans(windows, *host_index(t))
    ev = ops.EventStream.from_numpy(t, x, y, p)
    got = ops.event_volume_spans(ev, segments, spans, (H, W), K)
    assert float(got[0].abs().max()) == 0.0 and float(got[3].abs().max()) == 0.0
    for i in (1, 2):
        assert close(got[i], oracle_volume(t, x, y, p, windows[i], (H, W), K)), i
    # gen1 epilogue: nearest resize to 256 x 320, clamp, truncation
    want = torch.stack([ops.quantize_u8(ops.nearest_resize(v, (256, 320)), clamp255=True) for v in got])
    assert torch.equal(ops.event_volume_u8_batch(got, (256, 320)), want)


def test_driver_chunks_agree(tmp_path):
    """`generate_eventvolume.encode_recording` with different numbers of labels per call."""
    (t, x, y, p), labels = synth.write_recording(str(tmp_path), str(tmp_path), "train", "r", "gen1", 1_300_000, 3e5, 19)
    from frlw_evd_b200.recordings import DeviceRecording, Geometry
    rec = DeviceRecording(str(tmp_path / "train" / "r_td.dat"))
    geom = Geometry.for_dataset("gen1")
    a = [(lab, u8.cpu()) for lab, u8 in gev.encode_recording(rec, labels, geom, labels_per_call=4)]
    b = [(lab, u8.cpu()) for lab, u8 in gev.encode_recording(rec, labels, geom, labels_per_call=1000)]
    assert len(a) == len(b) > 0
    for (la, ua), (lb, ub) in zip(a, b):
        assert la == lb and torch.equal(ua, ub) and ua.shape == (3, 10, 256, 320)


def test_only_empty_windows():
    H, W, K = 32, 48, 5
    t = np.arange(0, 1000, dtype=np.uint32)
    z = np.zeros(1000, dtype=np.uint16)
    ev = ops.EventStream.from_numpy(t, z, z, z.astype(np.uint8))
    segments, spans = ops.plan_ev_spans([(5, 5, 0, 1000), (900, 900, 100, 50)], *host_index(t))
    assert segments == [] and spans == [(0, -1, 0, 1000), (0, -1, 100, 50)]
    u8 = torch.full((2, 2 * K, H, W), 7, dtype=torch.uint8, device=DEV)
    got = ops.event_volume_spans(ev, segments, spans, (H, W), K, out_u8=u8)
    assert float(got.abs().max()) == 0.0 and int(u8.max()) == 0
