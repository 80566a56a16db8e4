"""The whole-stream TAF path (bucketing + persistent tile kernel) against the oracle's
bin-by-bin execution of the same windows.  Float tolerance 1e-5 rel / 1e-6 abs."""
import numpy as np
import pytest
import torch

from frlw_evd_b200 import generate_taf as gt
from frlw_evd_b200 import ops, synth
from oracle import encoders as oe

from helpers import oracle_taf_windows

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6
DEV = "cuda"


def close(a, b):
    return np.allclose(a.cpu().numpy(), b.cpu().numpy() if torch.is_tensor(b) else b, rtol=RTOL, atol=ATOL)


def idx(t, v):
    return int(np.searchsorted(t, v))


@pytest.mark.parametrize("K", [8, 4])
def test_small_grid_mixed_windows(K):
    """Fresh + incremental windows, a long first window (> 32 bins), empty bins, a gap in
    the event range, a zero-bin window and an event exactly on a bin edge."""
    H, W, abin = 24, 40, 1000
    rng = np.random.Generator(np.random.PCG64(5))
    n = 30000
    t = np.sort(rng.integers(0, 120000, n)).astype(np.uint32)
    t[(t >= 50000) & (t < 53000)] = 49999            # three empty bins
    t[100] = 1000                                     # exactly on an edge: later bin wins
    t = np.sort(t)
    x = np.where(rng.random(n) < 0.4, rng.integers(0, 3, n), rng.integers(0, W, n)).astype(np.uint16)
    y = np.where(rng.random(n) < 0.4, rng.integers(0, 2, n), rng.integers(0, H, n)).astype(np.uint16)
    p = rng.integers(0, 2, n).astype(np.uint8)
    windows = [
        (0, idx(t, 60000), 0, 60, 1),                               # 60 bins: two batches
        (idx(t, 60000), idx(t, 65000), 60000, 5, 0),
        (idx(t, 65000), idx(t, 65000), 65000, 0, 0),                # zero bins: state re-emitted
        (idx(t, 65000), idx(t, 72500), 65000, 8, 0),                # last bin half filled
        (idx(t, 90000), idx(t, 98000), 90000, 8, 1),                # gap, then fresh
        (idx(t, 98000), idx(t, 104000), 98000, 6, 0),
    ]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    state = ops.taf_fresh_state((H, W), K, DEV)
    got = ops.taf_stream(ev, windows, abin, (H, W), K, state, emit_state_every_window=(K == 4))
    for i in range(len(windows)):
        assert close(got[i], want[i]), i
    assert close(state, want_state)
    # determinism: integer sums make the result independent of record order
    state2 = ops.taf_fresh_state((H, W), K, DEV)
    again = ops.taf_stream(ev, windows, abin, (H, W), K, state2)
    assert torch.equal(got, again) and torch.equal(state, state2)
    # continuing from a saved state == one launch over all windows
    s3 = ops.taf_fresh_state((H, W), K, DEV)
    a = ops.taf_stream(ev, windows[:2], abin, (H, W), K, s3)
    b = ops.taf_stream(ev, windows[2:4], abin, (H, W), K, s3)
    assert torch.equal(torch.cat([a, b]), got[:4])


def test_gen1_recording_plan_and_tensors(tmp_path):
    (t, x, y, p), labels = synth.write_recording(str(tmp_path), str(tmp_path), "train", "r", "gen1", 400000, 1e6, 1000)
    from frlw_evd_b200.recordings import DeviceRecording
    rec = DeviceRecording(str(tmp_path / "train" / "r_td.dat"))
    plan = gt.plan_windows(rec.loader, labels)
    windows = [w.as_tuple() for w in plan]
    want, _ = oracle_taf_windows(t, x, y, p, windows, 10000, (240, 304), 8)
    state = ops.taf_fresh_state((240, 304), 8, DEV)
    got = ops.taf_stream(rec.events, windows, 10000, (240, 304), 8, state)
    for i in range(len(windows)):
        assert close(got[i], want[i]), i


def test_gen4_policy_recording(tmp_path):
    (t, x, y, p), labels = synth.write_recording(str(tmp_path), str(tmp_path), "train", "r", "gen4", 200000, 5e6, 1002)
    from frlw_evd_b200.recordings import DeviceRecording, Geometry
    rec = DeviceRecording(str(tmp_path / "train" / "r_td.dat"))
    geom = Geometry.for_dataset("gen4")
    plan = gt.plan_windows(rec.loader, labels)
    windows = [w.as_tuple() for w in plan]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, 10000, (512, 640), 8, scale=(640 / 1280, 512 / 720))
    state = ops.taf_fresh_state((512, 640), 8, DEV)
    got = ops.taf_stream(rec.events, windows, 10000, (512, 640), 8, state, geom.coord_maps)
    for i in range(len(windows)):
        assert close(got[i], want[i]), i
    assert close(state, want_state)
    # stream path == bin-by-bin CUDA path (property that also holds at full benchmark size)
    s2 = ops.taf_fresh_state((512, 640), 8, DEV)
    w0 = windows[0]
    for b in range(w0[3]):
        lo, hi = idx(t, w0[2] + b * 10000), idx(t, w0[2] + (b + 1) * 10000)
        out, s2 = ops.taf_bin(rec.events.slice(max(lo, w0[0]), min(hi, w0[1])), w0[2] + b * 10000, 10000 + 1e-8,
                              (512, 640), 8, s2, geom.coord_maps)
    assert close(out, got[0])


def test_many_chunks_per_bucketing_cta_matches_bin_by_bin():
    """3 M events: more 4096-event chunks than bucketing CTAs (persistent loop), checked
    against the independent bin-by-bin CUDA path (oracle too slow at this size)."""
    t, x, y, p = synth.make_stream(720, 1280, 300000, 1e7, 77)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    windows, prev = [], 0
    for end in (100000, 150000, 200000, 250000, 300000):
        windows.append((idx(t, prev), idx(t, end), prev, (end - prev) // 10000, int(prev == 0)))
        prev = end
    state = ops.taf_fresh_state((512, 640), 8, DEV)
    got = ops.taf_stream(ev, windows, 10000, (512, 640), 8, state, maps)
    s2 = ops.taf_fresh_state((512, 640), 8, DEV)
    for b in range(30):
        lo, hi = idx(t, b * 10000), idx(t, (b + 1) * 10000)
        out, s2 = ops.taf_bin(ev.slice(lo, hi), b * 10000, 10000 + 1e-8, (512, 640), 8, s2, maps, in_place=True)
        if (b + 1) * 10000 in (100000, 150000, 200000, 250000, 300000):
            w = [100000, 150000, 200000, 250000, 300000].index((b + 1) * 10000)
            assert close(got[w], out), w
    assert close(state, s2)


def test_host_pipeline_matches_single_launch(tmp_path):
    """Chunked three-stream host pipeline == one launch + per-window epilogue."""
    (t, x, y, p), labels = synth.write_recording(str(tmp_path), str(tmp_path), "train", "r", "gen1", 500000, 1e6, 11)
    from frlw_evd_b200.recordings import DeviceRecording, Geometry
    rec = DeviceRecording(str(tmp_path / "train" / "r_td.dat"))
    geom = Geometry.for_dataset("gen1")
    plan = gt.plan_windows(rec.loader, labels)
    want = torch.stack([u8 for _, u8 in gt.encode_recording(rec, plan, geom)]).cpu()
    raw = torch.from_numpy(np.array(rec.loader.raw_bytes())).pin_memory()
    pipe = gt.HostPipeline(geom, plan, windows_per_chunk=3)
    out = torch.empty(pipe.out_shape, dtype=torch.uint8).pin_memory()
    pipe.run(raw, out)
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    pipe2 = gt.HostPipeline(geom, plan, windows_per_chunk=100)      # one chunk
    out2 = torch.empty(pipe2.out_shape, dtype=torch.uint8).pin_memory()
    pipe2.run(raw, out2)
    torch.cuda.synchronize()
    assert torch.equal(out2, want)


@pytest.mark.parametrize("K", [5, 8])
def test_event_volume_stream_matches_oracle_and_single_window_path(K):
    """Whole-stream Event Volume (bucketing + tile kernel) on the gen4 grid: oracle parity on
    a few windows, and agreement with the per-window CUDA path on all of them."""
    t, x, y, p = synth.make_stream(720, 1280, 400000, 5e6, 21)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    tw = 50000
    bounds = [0, 50000, 100000, 100000, 150000, 260000, 310000]       # one empty window, one gap
    windows = []
    for a in (0, 50000, 100000, 150000, 260000, 310000):
        windows.append((idx(t, a), idx(t, a + tw), a))
    windows.insert(3, (idx(t, 150000), idx(t, 150000), 150000))       # zero events
    windows.sort(key=lambda w: (w[0], w[1]))
    got = ops.event_volume_stream(ev, windows, tw, (512, 640), K, maps)
    for i, (lo, hi, t0) in enumerate(windows):
        single = ops.event_volume(ev.slice(lo, hi), t0, tw, (512, 640), K, maps)
        assert close(got[i], single), i
    from helpers import staged
    for i in (0, 4):
        lo, hi, t0 = windows[i]
        e = staged(t, x, y, p, lo, hi)
        e[:, 2] = (e[:, 2] - t0) / tw
        e[:, 0] *= 0.5
        e[:, 1] *= 512 / 720
        assert close(got[i], oe.event_volume(e, (512, 640), K)), i


def test_event_volume_stream_hot_pixels_and_determinism():
    """Cells that collect thousands of events (the fixed-point accumulator wraps several times)
    still match the oracle, and two runs are bit-identical."""
    rng = np.random.default_rng(5)
    n = 60000
    t = np.sort(rng.integers(0, 100000, n)).astype(np.uint32)
    x = rng.integers(0, 304, n).astype(np.uint16)
    y = rng.integers(0, 240, n).astype(np.uint16)
    p = rng.integers(0, 2, n).astype(np.uint8)
    hot = rng.random(n) < 0.5                       # half of the stream fires on three pixels
    which = rng.integers(0, 3, n)
    x[hot] = np.array([7, 150, 303], dtype=np.uint16)[which[hot]]
    y[hot] = np.array([0, 120, 239], dtype=np.uint16)[which[hot]]
    p[hot & (which == 1)] = 1
    ev = ops.EventStream.from_numpy(t, x, y, p)
    tw = 100000
    windows = [(0, n, 0)]
    got = ops.event_volume_stream(ev, windows, tw, (240, 304), 5).clone()
    again = ops.event_volume_stream(ev, windows, tw, (240, 304), 5)
    assert torch.equal(got, again)
    from helpers import staged
    e = staged(t, x, y, p, 0, n)
    e[:, 2] = (e[:, 2] - 0) / tw
    want = oe.event_volume(e, (240, 304), 5)
    assert float(want.max()) > 512 / 5 * 255                      # a cell sum beyond one wrap of the low word
    assert close(got[0], want)


def test_single_role_kernel_equals_warp_specialised(monkeypatch):
    """The non-specialised tile kernel (fallback when the staging tile does not fit next to two
    accumulator buffers) gives bit-identical results."""
    t, x, y, p = synth.make_stream(720, 1280, 150000, 8e6, 31)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    windows = [(0, idx(t, 80000), 0, 8, 1), (idx(t, 80000), idx(t, 150000), 80000, 7, 0)]
    outs = []
    for mode in ("ws", "single", "pk"):
        monkeypatch.setenv("EVREP_TAF_TILE_KERNEL", mode)
        state = ops.taf_fresh_state((512, 640), 8, DEV)
        outs.append((ops.taf_stream(ev, windows, 10000, (512, 640), 8, state, maps).clone(), state))
    for other in outs[1:]:
        assert torch.equal(outs[0][0], other[0]) and torch.equal(outs[0][1], other[1])


@pytest.mark.parametrize("K", [4, 8])
def test_packed_accumulator_kernel_hot_cells_and_dense_bins(K, monkeypatch):
    """EVREP_TAF_TILE_KERNEL=pk packs (sum d, n) of a cell into one word.  A pixel with more events in a bin than the
    sum field can hold exactly (> 419 at 10 ms), a bin with more records than one pass of the accumulate warps (768), more
    than the 10-bit count field (1023) and more than the record ring (8192): all bit-identical to the ws kernel and
    equal to the oracle."""
    H, W, abin = 48, 64, 10000
    rng = np.random.Generator(np.random.PCG64(77))
    parts = []
    def burst(n, t0, t1, hot=None):
        t = rng.integers(t0, t1, n)
        x = rng.integers(0, W, n); y = rng.integers(0, H, n); p = rng.integers(0, 2, n)
        if hot is not None:
            k = hot[0]
            x[:k], y[:k], p[:k] = hot[1], hot[2], hot[3]
        parts.append(np.stack([t, x, y, p], 1))
    burst(3000, 0, 10000, hot=(700, 5, 7, 1))           # bin 0: a pixel with 700 events, 3000 records in the tile
    burst(900, 10000, 20000)                            # bin 1: one pass + a bit
    burst(20000, 30000, 40000, hot=(5000, 60, 40, 0))   # bin 3 (bin 2 empty): beyond the ring, a very hot pixel
    burst(40, 40000, 50000)
    burst(1200, 70000, 80000, hot=(430, 0, 0, 0))       # second window
    e = np.concatenate(parts)
    e = e[np.argsort(e[:, 0], kind="stable")]
    t = e[:, 0].astype(np.uint32); x = e[:, 1].astype(np.uint16); y = e[:, 2].astype(np.uint16); p = e[:, 3].astype(np.uint8)
    cut = idx(t, 50000)
    windows = [(0, cut, 0, 5, 1), (cut, len(t), 50000, 3, 0)]
    ev = ops.EventStream.from_numpy(t, x, y, p)
    outs = []
    for mode in ("ws", "pk"):
        monkeypatch.setenv("EVREP_TAF_TILE_KERNEL", mode)
        state = ops.taf_fresh_state((H, W), K, DEV)
        outs.append((ops.taf_stream(ev, windows, abin, (H, W), K, state).clone(), state))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    assert close(outs[1][0][0], want[0]) and close(outs[1][0][1], want[1]) and close(outs[1][1], want_state)


def test_stream_on_unaligned_arrays_and_window_starts():
    """The bucketing passes use vector loads and TMA-staged chunks when the event arrays are 16-byte aligned and fall back
    to scalar loads otherwise; chunks start on a multiple of 16 events at or before the first window.  Same tensors from
    a stream that starts 3 events into its buffers, and for windows that begin at odd event indices."""
    H, W, K, abin = 240, 304, 8, 10000
    t, x, y, p = synth.make_stream(H, W, 60000, 2e6, 41)
    n = len(t)
    pad = 3
    tp_, xp_, yp_, pp_ = (np.concatenate([np.zeros(pad, a.dtype), a]) for a in (t, x, y, p))
    shifted = ops.EventStream.from_numpy(tp_, xp_, yp_, pp_).slice(pad, pad + n)
    aligned = ops.EventStream.from_numpy(t, x, y, p)
    a0, a1 = idx(t, 10000) + 5, idx(t, 30000)
    windows = [(a0, a1, 10000, 2, 1), (a1, idx(t, 60000), 30000, 3, 0)]
    outs = []
    for ev in (aligned, shifted):
        state = ops.taf_fresh_state((H, W), K, DEV)
        outs.append((ops.taf_stream(ev, windows, abin, (H, W), K, state).clone(), state))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    assert close(outs[0][0][0], want[0]) and close(outs[0][0][1], want[1]) and close(outs[0][1], want_state)


def test_stream_argument_errors():
    from frlw_evd_b200 import _lib
    t, x, y, p = synth.make_stream(240, 304, 20000, 1e6, 3)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    state = ops.taf_fresh_state((240, 304), 8, DEV)
    w = [(0, ev.n, 0, 2, 1)]
    with pytest.raises(_lib.EvrepError, match="invalid argument"):
        ops.taf_stream(ev, w, 10000, (240, 304), 5, ops.taf_fresh_state((240, 304), 5, DEV))      # K must be 4 or 8
    with pytest.raises(_lib.EvrepError, match="out of range"):
        ops.taf_stream(ev, w, 300000, (240, 304), 8, state)                                        # abin > 18 bits
    with pytest.raises(_lib.EvrepError, match="invalid argument"):
        ops.taf_stream(ev, [(0, ev.n, 0, 2, 1), (10, ev.n, 0, 2, 0)], 10000, (240, 304), 8, state)  # overlapping windows
    with pytest.raises(_lib.EvrepError, match="out of range"):
        ops.event_volume_stream(ev, [(0, ev.n, 0)], 300000, (240, 304), 5)
    assert ops.taf_stream(ev, [], 10000, (240, 304), 8, state).shape[0] == 0                       # no windows: no-op
