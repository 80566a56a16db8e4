"""data/sparse_ops.py twins (S1-S6) and event_queue_tensor (N1) on the GPU against the golden
vectors recorded from the reference (float sums: 1e-5 rel / 1e-6 abs)."""
import numpy as np
import pytest
import torch

from frlw_evd_b200 import event_representations
from frlw_evd_b200.data import sparse_ops as so

pytestmark = pytest.mark.gpu
DEV = "cuda"


def close(a, b):
    return np.allclose(a.cpu().numpy() if torch.is_tensor(a) else a, b, rtol=1e-5, atol=1e-6)


def G(golden, key):
    return torch.from_numpy(golden[key]).to(DEV)


def test_agile_event_volume_full_and_incremental(golden):
    H, W = [int(v) for v in golden["shape"]]
    v, st = so.generate_agile_event_volume_cuda(G(golden, "sp_events"), 2, (H, W), 0, None, 50000, 5, 10000)
    assert v.shape == golden["sp_agile_full"].shape and close(v, golden["sp_agile_full"])
    assert close(st, golden["sp_agile_full_state"])
    past = st.clone()
    v2, st2 = so.generate_agile_event_volume_cuda(G(golden, "sp_inc_events"), 2, (H, W), 60000, past, 50000, 5, 10000)
    assert close(v2, golden["sp_agile_inc"]) and close(st2, golden["sp_agile_inc_state"])
    # the caller's tensor was updated in place, like in the reference (sparse_ops.py:30-31)
    assert close(past[:, -1], golden["sp_agile_inc_state"][:, -2])


def test_event_volume_with_memory(golden):
    H, W = [int(v) for v in golden["shape"]]
    v, mem = so.generate_event_volume_cuda(G(golden, "sp_events"), 2, (H, W), 50000, None, 50000, 5, 10000)
    assert close(v, golden["sp_ev"]) and np.array_equal(mem.cpu().numpy(), golden["sp_ev_mem"])
    v2, mem2 = so.generate_event_volume_cuda(G(golden, "sp_inc_events")[:500], 2, (H, W), 60000, mem, 50000, 5, 10000)
    assert close(v2, golden["sp_ev2"]) and np.array_equal(mem2.cpu().numpy(), golden["sp_ev2_mem"])


def test_sparse_taf_frame_dense(golden):
    H, W = [int(v) for v in golden["shape"]]
    v, none = so.generate_taf_cuda(G(golden, "sp_taf_events"), 2, (H, W), 0, None, 50000, 5, 10000)
    assert none is None and close(v, golden["sp_taf"])
    f, _ = so.generate_event_frame_cuda(G(golden, "sp_events"), 2, (H, W), 0)
    assert np.array_equal(f.cpu().numpy(), golden["sp_frame"])
    dense = so.sparseToDense(G(golden, "sp_loc"), G(golden, "sp_feat"), (2, H, W))
    assert close(dense, golden["sp_dense"])
    loc, feat = so.denseToSparse(G(golden, "sp_dense"))
    assert np.array_equal(loc.cpu().numpy(), golden["sp_d2s_loc"]) and np.array_equal(feat.cpu().numpy(), golden["sp_d2s_feat"])


def test_event_queue_tensor(golden):
    H, W = [int(v) for v in golden["shape"]]
    got = event_representations.event_queue_tensor(golden["q_events"], 5, 2, H, W, golden["q_start"], 10000)
    assert got.dtype == np.float64 and got.shape == golden["q_out"].shape
    assert np.allclose(got, golden["q_out"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(got[1], golden["q_out"][1]) and np.array_equal(got[0, :4], golden["q_out"][0, :4])


def test_fetcher_streams_agile_event_volume_like_the_oracle():
    """``data/fetcher.py`` twin driving ``sparse_ops.generate_agile_event_volume_cuda`` on the
    device (events uploaded once, one index range per step) against the oracle encoder fed with
    the host-side boolean masks of the reference fetcher (``data/fetcher.py:35-55``)."""
    from frlw_evd_b200.data import fetcher as ft
    from oracle import encoders as oe
    rng = np.random.default_rng(8)
    n, B, H, W = 60000, 2, 30, 40
    ev = np.stack([rng.integers(0, B, n), rng.integers(0, W, n), rng.integers(0, H, n),
                   np.sort(rng.integers(0, 100000, n)), rng.integers(0, 2, n)], 1).astype(np.float64)
    labels = torch.tensor([[b, 1, 2, 3, 4, 0, 1000.0 * k] for b in range(B) for k in range(0, 200, 5)], dtype=torch.float64)
    timestamps = np.array([[0, 100000], [0, 100000]], dtype=np.int64)
    f = ft.fetcherVal(ev, (H, W), labels, timestamps, ["a", "b"], 50000, 5, 10000, so.generate_agile_event_volume_cuda)
    it, memory, steps = 0, None, 0
    while not f.finish:
        volume, _labels, _ts, _names, _secs = f.fetch()
        if it == 0:
            sel, it = ev[ev[:, 3] < 50000], 50000
        else:
            sel, it = ev[(ev[:, 3] < it + 10000) & (ev[:, 3] >= it)], it + 10000
        want, memory = oe.sparse_agile_event_volume(torch.from_numpy(sel), B, (H, W), it, memory, 50000, 5, 10000)
        assert f.iter == it and close(volume, want.numpy()), steps
        steps += 1
    assert steps == 6


def _batch_events(seed, n, B, H, W, t_max):
    rng = np.random.default_rng(seed)
    return np.stack([rng.integers(0, B, n), rng.integers(0, W, n), rng.integers(0, H, n),
                     np.sort(rng.integers(0, t_max, n)), rng.integers(0, 2, n)], 1).astype(np.float64)


def _reference_steps(ev, window, abin, total):
    """The host-side selections of the reference fetcher (data/fetcher.py:35-46): (events, clock after the step)."""
    it = 0
    while True:
        if it == 0:
            sel, it = ev[ev[:, 3] < window], window
        else:
            sel, it = ev[(ev[:, 3] < it + abin) & (ev[:, 3] >= it)], it + abin
        yield sel, it
        if it >= total:
            return


def test_fetcher_drives_event_volume_with_memory_frames_and_sparse_taf():
    """S2 (raw-event memory kept by the compaction kernel), S4 and S3 through the streaming fetcher, step by step
    against the oracle fed with the reference fetcher's host-side selections."""
    from frlw_evd_b200.data import fetcher as ft
    from oracle import encoders as oe
    B, H, W = 2, 30, 40
    ev = _batch_events(18, 50000, B, H, W, 90000)
    labels = torch.tensor([[b, 1, 2, 3, 4, 0, 1000.0 * k] for b in range(B) for k in range(0, 200, 5)], dtype=torch.float64)
    timestamps = np.array([[0, 90000], [0, 90000]], dtype=np.int64)
    for plugin, oracle in ((so.generate_event_volume_cuda, oe.sparse_event_volume), (so.generate_event_frame_cuda, oe.sparse_event_frame)):
        f = ft.fetcherVal(ev, (H, W), labels, timestamps, ["a", "b"], 50000, 5, 10000, plugin)
        memory, steps = None, 0
        for sel, it in _reference_steps(ev, 50000, 10000, 90000):
            volume, lab, _ts, _names, _secs = f.fetch()
            want, memory = oracle(torch.from_numpy(sel), B, (H, W), it, memory, 50000, 5, 10000)
            assert f.iter == it and close(volume, want.numpy()), (plugin.__name__, steps)
            if memory is not None:
                assert torch.equal(f.memory.cpu(), memory), steps           # the kept raw events, order included
            steps += 1
        assert f.finish and steps == 5 and len(f.represent_times()) == 5
    # S3: rows (b, x, y, t, c, p, feature)
    rng = np.random.default_rng(4)
    n = 20000
    ev7 = np.stack([rng.integers(0, B, n), rng.integers(0, W, n), rng.integers(0, H, n), np.sort(rng.integers(0, 60000, n)),
                    rng.integers(0, 10, n), rng.integers(0, 2, n), rng.random(n)], 1)
    f = ft.fetcherVal(ev7, (H, W), labels, np.array([[0, 60000], [0, 60000]]), ["a", "b"], 50000, 5, 10000, so.generate_taf_cuda)
    for sel, it in _reference_steps(ev7, 50000, 10000, 60000):
        volume, *_ = f.fetch()
        want, _ = oe.sparse_taf(torch.from_numpy(sel), B, (H, W), it, None, 50000, 5, 10000)
        assert close(volume, want.numpy())


def test_fetcher_online_taf_carries_the_state_on_the_device():
    """`generate_taf_online_cuda`: five bins on the first step, one per step afterwards, per sample the rule of
    generate_taf.py:19-58 (oracle: `taf_bin_update` bin by bin, sample by sample).  No step synchronises."""
    from frlw_evd_b200.data import fetcher as ft
    from oracle import encoders as oe
    B, H, W, K, abin = 3, 24, 32, 8, 10000
    ev = _batch_events(28, 40000, B, H, W, 80000)
    ev = ev[~((ev[:, 0] == 2) & (ev[:, 3] >= 60000) & (ev[:, 3] < 70000))]          # sample 2 is silent for one bin: it must not age
    labels = torch.tensor([[b, 1, 2, 3, 4, 0, 1000.0 * k] for b in range(B) for k in range(0, 200, 5)], dtype=torch.float64)
    timestamps = np.array([[0, 80000]] * B, dtype=np.int64)
    f = ft.fetcherVal(ev, (H, W), labels, timestamps, ["a", "b", "c"], 50000, K, abin, so.generate_taf_online_cuda)
    states = [oe.taf_fresh_state((H, W), K) for _ in range(B)]
    outs = [None] * B
    done = 0
    while not f.finish:
        volume, *_ = f.fetch()
        while done < f.iter:
            for b in range(B):
                rows = ev[(ev[:, 0] == b) & (ev[:, 3] >= done) & (ev[:, 3] < done + abin)]
                e5 = torch.from_numpy(np.stack([rows[:, 1], rows[:, 2], (rows[:, 3] - done) / (abin + 1e-8), rows[:, 4],
                                                np.zeros(len(rows))], 1))
                outs[b], states[b] = oe.taf_bin_update(e5, (H, W), states[b], K)
            done += abin
        for b in range(B):
            assert close(volume[b, ..., 0], outs[b].numpy()), (f.iter, b)
    assert f.iter == 80000
    for b in range(B):
        assert close(f.memory["state"][b], states[b].numpy()), b


def test_dense_to_sparse_kernel_matches_oracle_order_included():
    from oracle import encoders as oe
    rng = np.random.default_rng(6)
    dense = rng.standard_normal((3, 37, 53, 4)).astype(np.float32)
    dense[rng.random((3, 37, 53)) < 0.8] = 0.0
    dense[1, 5, 7] = [1.0, -1.0, 0.0, 0.0]                   # |.|-sum 2: kept although the plain sum is 0
    loc, feat = so.denseToSparse(torch.from_numpy(dense).to(DEV))
    want_loc, want_feat = oe.dense_to_sparse(torch.from_numpy(dense))
    assert torch.equal(loc.cpu(), want_loc) and torch.equal(feat.cpu(), want_feat)
    back = so.sparseToDense(torch.stack([loc[:, 2], loc[:, 0], loc[:, 1]], 1), feat, (3, 37, 53))
    assert torch.equal(back.cpu(), torch.from_numpy(dense))
    empty_loc, empty_feat = so.denseToSparse(torch.zeros((2, 8, 8, 3), device=DEV))
    assert empty_loc.shape == (0, 3) and empty_feat.shape == (0, 3)
