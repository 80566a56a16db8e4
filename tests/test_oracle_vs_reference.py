"""The oracle against the UNMODIFIED reference, run live (needs /root/reference: build container only;
skipped elsewhere -- the committed golden files carry the same pinning to the GPU box).  Fresh seeds
and sizes, so this is not a replay of tests/golden/: encoders bit for bit, one driver script byte for
byte."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import drivers as od
from oracle import encoders as oe
from oracle import ref_harness as rh

pytestmark = pytest.mark.reference


def events(seed, H, W, n, t_hi):
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.sort(rng.integers(0, t_hi, n))
    x = np.where(rng.random(n) < 0.25, rng.integers(0, 3, n), rng.integers(0, W, n))
    y = np.where(rng.random(n) < 0.25, rng.integers(0, 2, n), rng.integers(0, H, n))
    return np.stack([x, y, t, rng.integers(0, 2, n)], axis=1).astype(np.float64)


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_per_window_encoders_bit_for_bit():
    H, W, n = 20, 28, 5000
    ev = events(101, H, W, n, 40000)
    f = rh.load_functions("generate_eventcountimage.py")
    assert same(f["generate_eventframe"](torch.from_numpy(ev.copy()), (H, W))[0], oe.count_image(torch.from_numpy(ev.copy()), (H, W)))
    f = rh.load_functions("generate_eventvolume.py")
    evn = ev.copy()
    evn[:, 2] /= 40000
    for K in (5, 8):
        ref = f["generate_agile_event_volume_cuda"](torch.from_numpy(evn.copy()), (H, W), 40000, K)[0]
        assert same(ref, oe.event_volume(torch.from_numpy(evn.copy()), (H, W), K)), K
    torch.set_num_threads(1)                 # the reference's index_put_ on duplicate pixels races across CPU threads
    try:
        f = rh.load_functions("generate_surfaceofactiveevents.py")
        lam = [0.00001, 0.0000025, 0.000001]
        ref, ref_mem, _ = f["generate_leaky_cuda"](torch.from_numpy(ev[:3000].copy()), (H, W), lam, None, np.int64(30000))
        got, got_mem = oe.sae_surfaces(torch.from_numpy(ev[:3000].copy()), (H, W), lam, None, np.int64(30000))
        assert same(ref, got) and same(ref_mem, got_mem)
        ref2, ref_mem2, _ = f["generate_leaky_cuda"](torch.from_numpy(ev[3000:].copy()), (H, W), lam, ref_mem, np.int64(40000))
        got2, got_mem2 = oe.sae_surfaces(torch.from_numpy(ev[3000:].copy()), (H, W), lam, got_mem, np.int64(40000))
        assert same(ref2, got2) and same(ref_mem2, got_mem2)
    finally:
        torch.set_num_threads(os.cpu_count())


def test_taf_bins_and_leaky_bit_for_bit():
    H, W, K = 20, 28, 8
    ev = events(102, H, W, 5000, 50000)
    f = rh.load_functions("generate_taf.py")
    ref_state = torch.zeros((H, W, 2, K)) - 6000
    state = oe.taf_fresh_state((H, W), K)
    for it in range(5):
        sel = (ev[:, 2] >= it * 10000) & (ev[:, 2] < (it + 1) * 10000)
        e5 = np.concatenate([ev[sel], np.full((int(sel.sum()), 1), float(it))], axis=1)
        if it == 3:
            e5 = e5[:0]                       # an empty bin: no ageing
        e5[:, 2] = (e5[:, 2] - it * 10000) / (10000 + 1e-8)
        ref_out, ref_state, _ = f["generate_taf_cuda"](torch.from_numpy(e5.copy()), (H, W), ref_state, K)
        out, state = oe.taf_bin_update(torch.from_numpy(e5.copy()), (H, W), state, K)
        assert same(ref_out, out) and same(ref_state, state), it
    assert same(f["leaky_transform"](ref_out.view(K, 2, H, W)), oe.leaky_transform(out.view(K, 2, H, W)))


def test_online_encoders_bit_for_bit():
    H, W, n, B = 20, 28, 4000, 2
    ev = events(103, H, W, n, 50000)
    rng = np.random.Generator(np.random.PCG64(104))
    evb = np.concatenate([rng.integers(0, B, n)[:, None].astype(np.float64), ev], axis=1)
    so = rh.load_sparse_ops()
    T = lambda a: torch.from_numpy(a.copy())
    ref, ref_st = so.generate_agile_event_volume_cuda(T(evb), B, (H, W), 0, None, 50000, 5, 10000)
    got, got_st = oe.sparse_agile_event_volume(T(evb), B, (H, W), 0, None, 50000, 5, 10000)
    assert same(ref, got) and same(ref_st, got_st)
    ref, ref_mem = so.generate_event_volume_cuda(T(evb), B, (H, W), 50000, None, 50000, 5, 10000)
    got, got_mem = oe.sparse_event_volume(T(evb), B, (H, W), 50000, None, 50000, 5, 10000)
    assert same(ref, got) and same(ref_mem, got_mem)
    assert same(so.generate_event_frame_cuda(T(evb), B, (H, W), 0)[0], oe.sparse_event_frame(T(evb), B, (H, W), 0)[0])
    dense = so.sparseToDense(torch.from_numpy(np.stack([evb[:, 0], ev[:, 1], ev[:, 0]], 1).astype(np.int64)),
                             torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)), (B, H, W))
    for a, b in zip(so.denseToSparse(dense), oe.dense_to_sparse(dense)):
        assert same(a, b)


def digest_tree(root):
    out = {}
    for folder, _, names in os.walk(root):
        for name in names:
            with open(os.path.join(folder, name), "rb") as fh:
                out[os.path.relpath(os.path.join(folder, name), root)] = hashlib.sha256(fh.read()).hexdigest()
    return out


@pytest.mark.parametrize("rep,script,run", [("taf", "generate_taf.py", od.run_taf),
                                            ("event_volume", "generate_eventvolume.py", od.run_event_volume)])
def test_driver_script_byte_for_byte(tmp_path, rep, script, run):
    """The reference script run end to end (runpy under the CPU shims) and the oracle driver write the same files."""
    from frlw_evd_b200 import synth
    raw = str(tmp_path / "raw")
    synth.write_recording(raw, raw, "train", "r0", "gen1", 300000, 6e5, 2001)
    synth.write_recording(raw, raw, "test", "r1", "gen1", 250000, 6e5, 2002, header=False)
    ref_dir, mine_dir = str(tmp_path / "ref"), str(tmp_path / "mine")
    rh.run_script(script, ["-raw_dir", raw, "-label_dir", raw, "-target_dir", ref_dir, "-dataset", "gen1"])
    run(raw, raw, mine_dir, "gen1")
    ref, mine = digest_tree(ref_dir), digest_tree(mine_dir)
    assert ref and ref == mine
