"""Host-side logic on CPU: the PSEELoader mirror and the TAF window planner against the
oracle restatement (itself pinned to the reference), file formats, label parsing."""
import numpy as np
import pytest

from frlw_evd_b200 import generate_taf as gt
from frlw_evd_b200 import synth
from frlw_evd_b200.io import PSEELoader, dat_events_tools, npy_events_tools
from oracle import drivers as od
from oracle import psee_io


@pytest.fixture(scope="module")
def recording(tmp_path_factory):
    root = tmp_path_factory.mktemp("rec")
    (t, x, y, p), labels = synth.write_recording(str(root), str(root), "train", "a", "gen1", 600000, 1e6, 7)
    synth.write_recording(str(root), str(root), "val", "b", "gen1", 300000, 5e5, 8, header=False)
    return root, (t, x, y, p), labels


def test_header_and_decode(recording):
    root, (t, x, y, p), _ = recording
    with open(root / "train" / "a_td.dat", "rb") as fh:
        start, ev_type, ev_size, size = dat_events_tools.parse_header(fh)
    assert (ev_type, ev_size, size) == (0, 8, [240, 304]) and start > 0
    with open(root / "val" / "b_td.dat", "rb") as fh:
        assert dat_events_tools.parse_header(fh) == (0, 0, 8, [None, None])     # headerless
    ev = dat_events_tools.load_td_data(str(root / "train" / "a_td.dat"), 1000, 10)
    assert np.array_equal(ev["t"], t[10:1010]) and np.array_equal(ev["x"], x[10:1010])
    assert np.array_equal(ev["y"], y[10:1010]) and np.array_equal(ev["p"], p[10:1010])
    assert dat_events_tools.count_events(str(root / "train" / "a_td.dat")) == len(t)


def test_write_event_buffer_round_trip(tmp_path):
    t, x, y, p = synth.make_stream(240, 304, 10000, 1e6, 3)
    buf = np.empty(len(t), dtype=[("t", "u4"), ("x", "u2"), ("y", "u2"), ("p", "u1")])
    buf["t"], buf["x"], buf["y"], buf["p"] = t, x, y, p
    fh = dat_events_tools.write_header(str(tmp_path / "w_td.dat"), 240, 304)
    dat_events_tools.write_event_buffer(fh, buf)
    fh.close()
    back = PSEELoader(str(tmp_path / "w_td.dat")).load_n_events(len(t))
    for k in "txyp":
        assert np.array_equal(back[k], buf[k])


def test_loader_matches_oracle_on_random_operations(recording):
    root, (t, _, _, _), _ = recording
    path = str(root / "train" / "a_td.dat")
    a, b = psee_io.Loader(path), PSEELoader(path)
    rng = np.random.default_rng(0)
    mid = len(t) // 2
    for _ in range(1500):
        op = int(rng.integers(0, 4))
        if op == 0:
            q = int(rng.choice([rng.integers(-5, 700000), t[rng.integers(0, len(t))], t[mid], t[mid // 2]]))
            assert a.seek_time(q) == b.seek_time(q)
        elif op == 1:
            q = int(rng.integers(-3, len(t) + 3))
            a.seek_event(q), b.seek_event(q)
        elif op == 2:
            if a._cursor >= len(t):
                continue
            q = int(rng.integers(0, 250000))
            ra, rb = a.load_n_events(q), b.load_n_events(q)
            assert all(np.array_equal(ra[k], rb[k]) for k in "txyp")
        else:
            q = int(rng.integers(1, 300000))
            ra, rb = a.load_delta_t(q), b.load_delta_t(q)
            assert all(np.array_equal(ra[k], rb[k]) for k in "txyp")
        assert (a.current_time, a.done, a._cursor) == (b.current_time, b.done, b.position)
    with pytest.raises(ValueError):
        b.load_delta_t(0)


def test_label_times_new_and_legacy_names(tmp_path):
    times = np.arange(100000, 400000, 50000)
    synth.write_bbox_npy(str(tmp_path / "n_bbox.npy"), times)
    synth.write_bbox_npy(str(tmp_path / "l_bbox.npy"), times, legacy_names=True)
    for name in ("n_bbox.npy", "l_bbox.npy"):
        got = npy_events_tools.read_label_times(str(tmp_path / name))
        assert np.array_equal(got, times) and np.array_equal(got, psee_io.read_label_times(str(tmp_path / name)))


@pytest.mark.parametrize("labels", [
    np.arange(100000, 600000, 50000),
    np.array([90000, 95000, 97000, 250000, 251000, 420000, 599999, 700000]),   # snapping, zero bins, gaps, past EOF
])
def test_taf_window_plan_matches_oracle(recording, labels):
    root, _, _ = recording
    path = str(root / "train" / "a_td.dat")
    plan = gt.plan_windows(PSEELoader(path), labels, min_event_count=200000)
    loader = psee_io.Loader(path)
    t_upper, c_upper, want = -1e16, -1, []
    for label in labels:
        w = od.taf_window_plan(loader, label, t_upper, c_upper, 10000, 80000, 200000)
        if w is None:
            continue
        want.append(w)
        t_upper, c_upper = w[2], w[4]
    got = [(w.fresh, w.start_time, w.end_time, w.start_count, w.end_count) for w in plan]
    assert got == [tuple(w) for w in want] and len(got) >= 5
    assert any(not w.fresh for w in plan) and sum(w.fresh for w in plan) >= 1


def test_count_stream_segment_plan():
    """``ops.plan_count_segments``: nested and overlapping windows become runs of consecutive
    segments; empty windows are left out; the emission order is by last segment."""
    from frlw_evd_b200 import ops
    windows = [(0, 0), (10, 30), (0, 30), (25, 60), (40, 60), (60, 60), (5, 100)]
    segments, order, runs = ops.plan_count_segments(windows)
    assert segments == [(0, 5), (5, 10), (10, 25), (25, 30), (30, 40), (40, 60), (60, 100)]
    assert all(a[1] == b[0] for a, b in zip(segments, segments[1:]))
    assert order == [1, 2, 3, 4, 6]
    for i, (first, last) in zip(order, runs):
        assert segments[first][0] == windows[i][0] and segments[last][1] == windows[i][1]
    assert [last for _, last in runs] == sorted(last for _, last in runs)
    assert ops.plan_count_segments([(3, 3)]) == ([], [], [])


def test_sae_and_count_stream_drivers_plan_like_the_oracle_drivers(tmp_path):
    """The window lists of the whole-stream SAE driver equal the event ranges the oracle driver
    encodes label by label."""
    import numpy as np
    from frlw_evd_b200 import generate_surfaceofactiveevents as g_sae, synth
    from frlw_evd_b200.io import PSEELoader
    from oracle.psee_io import Loader
    t, x, y, p = synth.make_stream(240, 304, 7_000_000, 2e4, 12)
    path = str(tmp_path / "rec_td.dat")
    synth.write_dat(path, t, x, y, p, 240, 304)
    labels = [300_000, 350_000, 5_900_000, 6_000_000, 6_050_000, 6_999_000, 8_000_000]
    plan = g_sae.plan_windows(PSEELoader(path), labels)
    loader, want = Loader(path), []
    t_upper, c_upper = -100000000, 0
    for label in labels:                       # oracle/drivers.py:run_sae, largest sub-window
        end_count = loader.seek_time(label)
        if end_count is None:
            continue
        start_time = label - 5000000
        start_count = loader.seek_time(0 if start_time < 0 else start_time)
        if start_count is None or start_time < 0:
            start_count = 0
        if start_time <= t_upper:
            start_count = c_upper
        t_upper, c_upper = label, end_count
        lo = start_count + int(np.searchsorted(t[start_count:end_count], label - 5541263, side="right"))
        want.append((label, lo, end_count))
    assert plan == want and len(plan) >= 5


def test_bind_to_device_is_best_effort():
    """``affinity.bind_to_device`` reports what it did and never raises (no NVML on the CPU box)."""
    import os
    from frlw_evd_b200.affinity import bind_to_device
    before = os.sched_getaffinity(0)
    info = bind_to_device(0)
    assert set(info) >= {"bound", "cpus", "how"}
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


@pytest.mark.reference
def test_fetcher_steps_like_the_reference_fetcher():
    """``data/fetcher.py`` twin against the unmodified reference class on CPU: the same events
    reach ``to_volume`` at every step, with the same ``iter`` / window arguments, and the labels,
    timestamps and finish flag agree (``data/fetcher.py:35-62``)."""
    import importlib.util
    import numpy as np
    import torch
    from oracle import ref_harness as rh
    from frlw_evd_b200.data import fetcher as mine
    spec = importlib.util.spec_from_file_location("ref_fetcher", rh.REFERENCE_ROOT + "/data/fetcher.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(2)
    n, B = 5000, 2
    ev = np.stack([rng.integers(0, B, n), rng.integers(0, 40, n), rng.integers(0, 30, n),
                   np.sort(rng.integers(0, 90000, n)), rng.integers(0, 2, n)], 1).astype(np.float64)
    labels = torch.tensor([[b, 1, 2, 3, 4, 0, 1000.0 * k + 50000] for b in range(B) for k in range(0, 45, 5)], dtype=torch.float64)
    timestamps = np.array([[1000, 91000], [2000, 92000]], dtype=np.int64)
    seen = {"ref": [], "mine": []}

    def recorder(tag):
        def to_volume(events, batch, shape, it, memory, window, bins, abin):
            seen[tag].append((np.asarray(events.cpu()), batch, tuple(shape), it, memory, window, bins, abin))
            return torch.zeros(1), len(seen[tag])
        return to_volume
    with rh._cpu_cuda_shims():
        a = ref.fetcherTrain(ev, (30, 40), labels, timestamps, ["x", "y"], 50000, 5, 10000, recorder("ref"))
        b = mine.fetcherTrain(ev, (30, 40), labels, timestamps, ["x", "y"], 50000, 5, 10000, recorder("mine"), device="cpu")
        for _ in range(5):
            ra, rb = a.fetch(), b.fetch()
            assert (ra[1] is None) == (rb[1] is None)
            if ra[1] is not None:
                assert torch.equal(ra[1], rb[1])
            assert np.array_equal(ra[2], rb[2]) and ra[3] == rb[3] and a.finish == b.finish and a.iter == b.iter
            if a.finish:
                break
    assert len(seen["ref"]) == len(seen["mine"]) >= 4
    for x, y in zip(seen["ref"], seen["mine"]):
        assert np.array_equal(x[0], y[0]) and x[1:] == y[1:]
