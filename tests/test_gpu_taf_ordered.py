"""The two TAF paths for time-ordered streams -- the default one-pass bin-major sort in front of the register-resident
tile kernel (`evrep_taf_stream_ordered`) and the slice sort + shared-memory tile kernel (`evrep_taf_stream_sliced`,
EVREP_TAF_PATH=sliced) -- against the oracle and against the general two-pass path (EVREP_TAF_PATH=bucketed): hot pixels (exact two-word
accumulators), wide offsets, bins of many slices, the native 1MP grid (several waves of tiles),
K = 4 under the gen4 policy, the in-kernel uint8 output and the order check.
Float tolerance 1e-5 rel / 1e-6 abs (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from frlw_evd_b200 import ops, synth

from helpers import oracle_taf_windows

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6
DEV = "cuda"


def close(a, b):
    return np.allclose(a.cpu().numpy(), b.cpu().numpy() if torch.is_tensor(b) else b, rtol=RTOL, atol=ATOL)


def idx(t, v):
    return int(np.searchsorted(t, v))


def run_path(monkeypatch, path, ev, windows, abin, grid, K, maps=None):
    monkeypatch.setenv("EVREP_TAF_PATH", path)
    state = ops.taf_fresh_state(grid, K, DEV)
    out = ops.taf_stream(ev, windows, abin, grid, K, state, maps)
    if path != "bucketed":
        assert ops.order_violations(DEV) == 0
    monkeypatch.delenv("EVREP_TAF_PATH", raising=False)
    return out, state


def both_paths(monkeypatch, ev, windows, abin, grid, K, maps=None):
    """(sliced result, sliced state, bucketed result, bucketed state); the default ordered path must agree with the
    bucketed one bit for bit (same tile kernel arithmetic, integer sums)."""
    a, s1 = run_path(monkeypatch, "sliced", ev, windows, abin, grid, K, maps)
    b, s2 = run_path(monkeypatch, "bucketed", ev, windows, abin, grid, K, maps)
    c, s3 = run_path(monkeypatch, "ordered", ev, windows, abin, grid, K, maps)
    assert torch.equal(b, c) and torch.equal(s2, s3)
    return a, s1, b, s2


def test_hot_pixel_takes_the_exact_path(monkeypatch):
    """20000 events on one cell inside one bin (and 300 on another: above the 256-count trigger of
    the packed accumulators) next to ordinary traffic."""
    H, W, K, abin = 48, 64, 8, 10000
    rng = np.random.Generator(np.random.PCG64(9))
    n_bg = 40000
    t = rng.integers(0, 60000, n_bg)
    x = rng.integers(0, W, n_bg); y = rng.integers(0, H, n_bg); p = rng.integers(0, 2, n_bg)
    hot_t = rng.integers(20000, 30000, 20000)
    warm_t = rng.integers(40000, 50000, 300)
    t = np.concatenate([t, hot_t, warm_t]); x = np.concatenate([x, np.full(20000, 17), np.full(300, 5)])
    y = np.concatenate([y, np.full(20000, 33), np.full(300, 6)]); p = np.concatenate([p, np.ones(20000, int), np.zeros(300, int)])
    order = np.argsort(t, kind="stable")
    t, x, y, p = t[order].astype(np.uint32), x[order].astype(np.uint16), y[order].astype(np.uint16), p[order].astype(np.uint8)
    windows = [(0, idx(t, 30000), 0, 3, 1), (idx(t, 30000), len(t), 30000, 3, 0)]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    got, state, old, _ = both_paths(monkeypatch, ev, windows, abin, (H, W), K)
    for i in range(2):
        assert close(got[i], want[i]), i
        assert close(old[i], want[i]), i                 # the default path on the same hot pixel
    assert close(state, want_state)
    # the hot cell itself: mean of 20000 offsets; the reference sums float32 sequentially, the kernel sums integers
    cell = got[0][2 * (K - 1) + 1, 33, 17].item()
    ref = want[0][2 * (K - 1) + 1, 33, 17].item()
    rel = abs(cell - ref) / abs(ref)
    print("hot cell: got %.9f reference %.9f rel err %.2e" % (cell, ref, rel))
    assert rel < 1e-5
    assert close(got, old)


def test_wide_offsets_and_clamped_edges(monkeypatch):
    """abin = 50 ms (offsets beyond the packed accumulators' 32767) against the oracle; events before the
    first and after the last bin edge (clamped) against the general path, which defines them."""
    H, W, K = 24, 40, 4
    rng = np.random.Generator(np.random.PCG64(3))
    n = 50000
    t = np.sort(rng.integers(0, 400000, n)).astype(np.uint32)
    x = rng.integers(0, W, n).astype(np.uint16); y = rng.integers(0, H, n).astype(np.uint16); p = rng.integers(0, 2, n).astype(np.uint8)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    windows = [(0, idx(t, 200000), 0, 4, 1), (idx(t, 200000), idx(t, 400000), 200000, 4, 0)]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, 50000, (H, W), K)
    got, state, old, _ = both_paths(monkeypatch, ev, windows, 50000, (H, W), K)
    for i in range(2):
        assert close(got[i], want[i]), i
    assert close(state, want_state) and close(got, old)
    # window [100 ms, 130 ms) in 10 ms bins over events from 80 ms to 170 ms: 20 ms of events before bin 0, 40 ms after the last
    clamped = [(idx(t, 80000), idx(t, 170000), 100000, 3, 1)]
    got, state, old, old_state = both_paths(monkeypatch, ev, clamped, 10000, (H, W), K)
    assert close(got, old) and close(state, old_state)


def test_bin_of_many_slices(monkeypatch):
    """400 k events inside one 10 ms bin: 49 slices of 8188, i.e. more than one 32-slice batch of the
    producer warp, and stages that fill up in the middle of a run."""
    H, W, K, abin = 96, 128, 8, 10000
    rng = np.random.Generator(np.random.PCG64(21))
    n = 500000
    t = np.sort(np.concatenate([rng.integers(0, 10000, 400000), rng.integers(10000, 40000, 100000)])).astype(np.uint32)
    x = rng.integers(0, W, n).astype(np.uint16); y = rng.integers(0, H, n).astype(np.uint16); p = rng.integers(0, 2, n).astype(np.uint8)
    windows = [(0, idx(t, 20000), 0, 2, 1), (idx(t, 20000), n, 20000, 2, 0)]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    got, state, _, _ = both_paths(monkeypatch, ev, windows, abin, (H, W), K)
    for i in range(2):
        assert close(got[i], want[i]), i
    assert close(state, want_state)


@pytest.fixture
def ordered_path(monkeypatch):
    monkeypatch.setenv("EVREP_TAF_PATH", "sliced")


@pytest.mark.parametrize("path", ["ordered", "sliced", "bucketed"])
@pytest.mark.parametrize("K", [8, 4])
def test_native_1mp_grid_against_oracle(K, path, monkeypatch):
    """720 x 1280 without down-scaling: more tiles than resident CTAs (several waves), a fresh window
    and two incremental ones."""
    monkeypatch.setenv("EVREP_TAF_PATH", path)
    H, W, abin = 720, 1280, 10000
    t, x, y, p = synth.make_stream(H, W, 70000, 2e7, 31)
    windows = [(0, idx(t, 30000), 0, 3, 1), (idx(t, 30000), idx(t, 50000), 30000, 2, 0), (idx(t, 50000), idx(t, 70000), 50000, 2, 0)]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    assert ev.is_ordered()
    state = ops.taf_fresh_state((H, W), K, DEV)
    got = ops.taf_stream(ev, windows, abin, (H, W), K, state)
    for i in range(3):
        assert close(got[i], want[i]), i
    assert close(state, want_state)


@pytest.mark.parametrize("path", ["ordered", "sliced", "bucketed"])
def test_k4_gen4_policy_against_oracle(path, monkeypatch):
    """K = 4 on the down-scaled 512 x 640 grid (gen4 policy: float64 scale, truncation)."""
    monkeypatch.setenv("EVREP_TAF_PATH", path)
    t, x, y, p = synth.make_stream(720, 1280, 80000, 1e7, 41)
    windows = [(0, idx(t, 40000), 0, 4, 1), (idx(t, 40000), idx(t, 80000), 40000, 4, 0)]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, 10000, (512, 640), 4, scale=(640 / 1280, 512 / 720))
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    state = ops.taf_fresh_state((512, 640), 4, DEV)
    got = ops.taf_stream(ev, windows, 10000, (512, 640), 4, state, maps)
    for i in range(2):
        assert close(got[i], want[i]), i
    assert close(state, want_state)


def test_uint8_straight_from_the_tile_kernel(ordered_path):
    """out_u8 against the fused epilogue applied to the float tensors of the same call (within one step on
    fewer than 0.1 % of the bytes); also without any float output at all."""
    t, x, y, p = synth.make_stream(720, 1280, 100000, 1e7, 51)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    K, grid = 8, (512, 640)
    windows = [(idx(t, a), idx(t, a + 50000), a, 5, int(a == 0)) for a in (0, 50000)]
    state = ops.taf_fresh_state(grid, K, DEV)
    u8 = torch.zeros((2, K, 2, 512, 640), dtype=torch.uint8, device=DEV)
    vol = ops.taf_stream(ev, windows, 10000, grid, K, state, maps, out_u8=u8)
    # the tile kernel takes the logarithm from MUFU.LG2, the epilogue kernel calls log1pf: values that sit on an integer
    # boundary before truncation may differ by one step
    diff = (u8.int() - ops.taf_leaky_u8_batch(vol, K).int()).abs()
    assert int(diff.max()) <= 1 and float((diff != 0).float().mean()) < 1e-3
    state2 = ops.taf_fresh_state(grid, K, DEV)
    u8b = torch.zeros_like(u8)
    assert ops.taf_stream(ev, windows, 10000, grid, K, state2, maps, out_u8=u8b, want_f32=False) is None
    assert torch.equal(u8, u8b) and torch.equal(state, state2)


@pytest.mark.parametrize("path", ["ordered", "sliced"])
def test_unordered_input_is_detected_and_routed(path, monkeypatch):
    """A stream with a few timestamps out of order: `is_ordered` sends it to the general path (oracle
    parity); forcing it through an ordered entry point reports the violations."""
    monkeypatch.setenv("EVREP_TAF_PATH", path)
    H, W, K, abin = 24, 40, 8, 1000
    rng = np.random.Generator(np.random.PCG64(13))
    n = 20000
    t = np.sort(rng.integers(0, 20000, n)).astype(np.uint32)
    swap = rng.choice(n - 2000, 50, replace=False)
    t[swap], t[swap + 1500] = t[swap + 1500].copy(), t[swap].copy()      # ~1.5 bins out of place
    x = rng.integers(0, W, n).astype(np.uint16); y = rng.integers(0, H, n).astype(np.uint16); p = rng.integers(0, 2, n).astype(np.uint8)
    windows = [(0, n, 0, 20, 1)]
    want, _ = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    assert not ev.is_ordered()
    state = ops.taf_fresh_state((H, W), K, DEV)
    got = ops.taf_stream(ev, windows, abin, (H, W), K, state)
    assert close(got[0], want[0])
    ev.ordered = True                                                     # lie: the ordered path must notice
    ops.taf_stream(ev, windows, abin, (H, W), K, ops.taf_fresh_state((H, W), K, DEV))
    assert ops.order_violations(DEV) > 0


def test_large_windows_fall_back_to_the_general_path(monkeypatch):
    """A window of more than 128 slices is refused by the one-pass path: `taf_stream` routes it to the two-pass one."""
    monkeypatch.setenv("EVREP_TAF_PATH", "ordered")
    H, W, K, abin = 32, 48, 8, 10000
    rng = np.random.Generator(np.random.PCG64(23))
    n = ops.TAF_ORDERED_MAX_WINDOW + 5000
    t = np.sort(rng.integers(0, 20000, n)).astype(np.uint32)
    x = rng.integers(0, W, n).astype(np.uint16); y = rng.integers(0, H, n).astype(np.uint16); p = rng.integers(0, 2, n).astype(np.uint8)
    windows = [(0, n, 0, 2, 1)]
    ev = ops.EventStream.from_numpy(t, x, y, p)
    state = ops.taf_fresh_state((H, W), K, DEV)
    got = ops.taf_stream(ev, windows, abin, (H, W), K, state)
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    assert close(got[0], want[0]) and close(state, want_state)
    assert int(ops.order_violations_tensor(DEV).abs().sum()) == 0          # no ordered call was made


def test_long_window_rebases_inside(ordered_path):
    """One fresh window of 100 non-empty bins: the lazy ageing counter is folded back every 8 bins."""
    H, W, K, abin = 24, 40, 8, 1000
    rng = np.random.Generator(np.random.PCG64(17))
    n = 60000
    t = np.sort(rng.integers(0, 100000, n)).astype(np.uint32)
    x = rng.integers(0, W, n).astype(np.uint16); y = rng.integers(0, H, n).astype(np.uint16); p = rng.integers(0, 2, n).astype(np.uint8)
    keep = (x > 3) | (t < 5000)                                           # a few columns fall silent early and only age
    t, x, y, p = t[keep], x[keep], y[keep], p[keep]
    windows = [(0, len(t), 0, 100, 1)]
    want, want_state = oracle_taf_windows(t, x, y, p, windows, abin, (H, W), K)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    state = ops.taf_fresh_state((H, W), K, DEV)
    got = ops.taf_stream(ev, windows, abin, (H, W), K, state)
    assert close(got[0], want[0]) and close(state, want_state)


@pytest.mark.parametrize("path", ["ordered", "sliced"])
def test_degenerate_inputs(path, monkeypatch):
    """Windows without bins, windows without events, an empty stream: the state is emitted unchanged."""
    monkeypatch.setenv("EVREP_TAF_PATH", path)
    H, W, K = 16, 24, 8
    t = np.arange(0, 5000, 5, dtype=np.uint32)
    n = len(t)
    x = (np.arange(n) % W).astype(np.uint16); y = (np.arange(n) % H).astype(np.uint16); p = (np.arange(n) % 2).astype(np.uint8)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    state = ops.taf_fresh_state((H, W), K, DEV)
    state += torch.arange(K, device=DEV, dtype=torch.float32)                # something recognisable
    before = state.clone()
    out = ops.taf_stream(ev, [(0, 0, 0, 0, 0), (0, 0, 0, 0, 0)], 1000, (H, W), K, state)       # no bins at all
    want = before.permute(3, 2, 0, 1).reshape(2 * K, H, W)
    assert torch.equal(out[0], want) and torch.equal(out[1], want) and torch.equal(state, before)
    out = ops.taf_stream(ev, [(n, n, 9000, 3, 0)], 1000, (H, W), K, state)                     # bins, but no events in them
    assert torch.equal(out[0], want) and torch.equal(state, before)
    empty = ops.EventStream.from_numpy(t[:0], x[:0], y[:0], p[:0])
    out = ops.taf_stream(empty, [(0, 0, 0, 2, 1)], 1000, (H, W), K, state)                     # fresh window on an empty stream
    assert torch.equal(out[0], torch.full((2 * K, H, W), -6000.0, device=DEV))
