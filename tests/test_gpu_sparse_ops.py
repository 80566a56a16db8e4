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
