"""Generate ``tests/golden/`` from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Run as ``python -m oracle.make_golden`` from the repo root while
``/root/reference`` is mounted.  Inputs are small seeded tensors; outputs are whatever
the reference's own functions / scripts / compiled extension produce on CPU
(``oracle/ref_harness.py``; SAE with one torch thread, because the reference's
``index_put_`` on duplicate indices races across CPU threads).  The committed files pin
the in-repo restatement (``tests/test_oracle_golden.py``) and, through it, the CUDA path.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from frlw_evd_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# driver fixtures: (tag, sensor, dataset flag, duration us, rate ev/s, [(mode, name, seed, header)])
DRIVER_CASES = [
    ("gen1", "gen1", "gen1", 400000, 1.0e6, [("train", "rec0", 1000, True), ("test", "rec1", 1001, False)]),
    ("gen4", "gen4", "gen4", 250000, 2.0e6, [("train", "rec0", 1002, True), ("test", "rec1", 1003, False)]),
]
DRIVER_SCRIPTS = {
    "count_image": "generate_eventcountimage.py",
    "sae": "generate_surfaceofactiveevents.py",
    "event_volume": "generate_eventvolume.py",
    "taf": "generate_taf.py",
}


def small_events(seed, H, W, n, t_hi):
    """Seeded small-grid events with many duplicate pixels and tied timestamps."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.sort(rng.integers(0, t_hi, n))
    x = np.where(rng.random(n) < 0.3, rng.integers(0, 4, n), rng.integers(0, W, n))
    y = np.where(rng.random(n) < 0.3, rng.integers(0, 3, n), rng.integers(0, H, n))
    p = rng.integers(0, 2, n)
    return np.stack([x, y, t, p], axis=1).astype(np.float64)


def encoder_golden():
    out = {}
    H, W, n = 24, 40, 6000
    ev = small_events(11, H, W, n, 50000)
    out["events"] = ev
    out["shape"] = np.array([H, W])

    f = rh.load_functions("generate_eventcountimage.py")
    out["eci"] = f["generate_eventframe"](torch.from_numpy(ev.copy()), (H, W))[0].numpy()

    f = rh.load_functions("generate_eventvolume.py")
    evn = ev.copy()
    evn[:, 2] = evn[:, 2] / 50000
    for K in (5, 8):
        out["ev_k%d" % K] = f["generate_agile_event_volume_cuda"](torch.from_numpy(evn.copy()), (H, W), 50000, K)[0].numpy()
    out["events_norm"] = evn

    torch.set_num_threads(1)
    f = rh.load_functions("generate_surfaceofactiveevents.py")
    lam = [0.00001, 0.0000025, 0.000001]
    ev_oob = ev.copy()
    ev_oob[::97, 0] = W + 3          # rows the SAE bounds filter must drop
    a, mem, _ = f["generate_leaky_cuda"](torch.from_numpy(ev_oob[:4000].copy()), (H, W), lam, None, np.int64(40000))
    out["sae_events"] = ev_oob
    out["sae_out0"], out["sae_mem0"] = a.numpy(), mem.numpy().copy()
    a, mem, _ = f["generate_leaky_cuda"](torch.from_numpy(ev_oob[4000:].copy()), (H, W), lam, mem, np.int64(50000))
    out["sae_out1"], out["sae_mem1"] = a.numpy(), mem.numpy().copy()
    torch.set_num_threads(os.cpu_count())

    f = rh.load_functions("generate_taf.py")
    K = 8
    state = torch.zeros((H, W, 2, K)) - 6000
    outs, states = [], []
    for it in range(6):
        sel = (ev[:, 2] >= it * 10000) & (ev[:, 2] < (it + 1) * 10000)
        e5 = np.concatenate([ev[sel], np.full((int(sel.sum()), 1), float(it))], axis=1)
        if it in (2, 5):
            e5 = e5[:0]               # empty bins: no ageing
        e5[:, 2] = (e5[:, 2] - it * 10000) / (10000 + 1e-8)
        o, state, _ = f["generate_taf_cuda"](torch.from_numpy(e5.copy()), (H, W), state, K)
        outs.append(o.numpy().copy())
        states.append(state.numpy().copy())
    out["taf_out"], out["taf_state"] = np.stack(outs), np.stack(states)
    out["taf_leaky"] = f["leaky_transform"](torch.from_numpy(outs[-2]).view(K, 2, H, W)).numpy()

    # data/sparse_ops.py (S1-S6)
    so = rh.load_sparse_ops()
    B = 2
    rng = np.random.Generator(np.random.PCG64(12))
    b = rng.integers(0, B, n)
    evb = np.concatenate([b[:, None].astype(np.float64), ev], axis=1)      # (b, x, y, t, p)
    out["sp_events"] = evb
    v, st = so.generate_agile_event_volume_cuda(torch.from_numpy(evb.copy()), B, (H, W), 0, None, 50000, 5, 10000)
    out["sp_agile_full"], out["sp_agile_full_state"] = v.numpy(), st.numpy().copy()
    inc = evb.copy()
    inc[:, 3] = 50000 + inc[:, 3] / 5     # events of the next 10 ms step
    v, st2 = so.generate_agile_event_volume_cuda(torch.from_numpy(inc.copy()), B, (H, W), 60000, st.clone(), 50000, 5, 10000)
    out["sp_inc_events"] = inc
    out["sp_agile_inc"], out["sp_agile_inc_state"] = v.numpy(), st2.numpy().copy()
    v, mem = so.generate_event_volume_cuda(torch.from_numpy(evb.copy()), B, (H, W), 50000, None, 50000, 5, 10000)
    out["sp_ev"], out["sp_ev_mem"] = v.numpy(), mem.numpy().copy()
    v2, mem2 = so.generate_event_volume_cuda(torch.from_numpy(inc[:500].copy()), B, (H, W), 60000, mem, 50000, 5, 10000)
    out["sp_ev2"], out["sp_ev2_mem"] = v2.numpy(), mem2.numpy().copy()
    c = rng.integers(0, 10, n).astype(np.float64)
    feat = rng.normal(size=n)
    ev7 = np.stack([evb[:, 0], evb[:, 1], evb[:, 2], evb[:, 3], c, evb[:, 4], feat], axis=1)
    out["sp_taf_events"] = ev7
    out["sp_taf"] = so.generate_taf_cuda(torch.from_numpy(ev7.copy()), B, (H, W), 0, None, 50000, 5, 10000)[0].numpy()
    out["sp_frame"] = so.generate_event_frame_cuda(torch.from_numpy(evb.copy()), B, (H, W), 0)[0].numpy()
    loc = np.stack([b, ev[:, 1], ev[:, 0]], axis=1).astype(np.int64)
    feats = rng.normal(size=(n, 3)).astype(np.float32)
    dense = so.sparseToDense(torch.from_numpy(loc), torch.from_numpy(feats), (B, H, W))
    out["sp_loc"], out["sp_feat"], out["sp_dense"] = loc, feats, dense.numpy()
    l2, f2 = so.denseToSparse(dense)
    out["sp_d2s_loc"], out["sp_d2s_feat"] = l2.numpy(), f2.numpy()

    # compiled reference extension (N1)
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import event_representations as er
    Q, abin = 5, 10000
    start = np.array([1000, 5000], dtype=np.int32)
    z = np.sort(rng.integers(0, 6, n))
    tq = start[b] + z * abin + rng.integers(0, abin, n)
    evq = np.stack([b, ev[:, 0], ev[:, 1], tq, ev[:, 3], z], axis=1).astype(np.float32)
    out["q_events"], out["q_start"] = evq, start
    out["q_out"] = er.event_queue_tensor(evq, Q, B, H, W, start, abin)

    np.savez_compressed(os.path.join(GOLDEN, "encoders_small.npz"), **out)
    return out


def _digest_tree(root):
    digests = {}
    for d, _, files in os.walk(root):
        for name in files:
            path = os.path.join(d, name)
            with open(path, "rb") as fh:
                digests[os.path.relpath(path, root)] = hashlib.sha256(fh.read()).hexdigest()
    return dict(sorted(digests.items()))


def write_case(raw, case):
    _tag, sensor, _flag, duration, rate, recs = case
    for mode, name, seed, header in recs:
        synth.write_recording(raw, raw, mode, name, sensor, duration, rate, seed, header=header)


def driver_golden():
    result = {}
    torch.set_num_threads(1)         # determinism of the reference's SAE scatter
    for case in DRIVER_CASES:
        tag, _sensor, flag = case[0], case[1], case[2]
        with tempfile.TemporaryDirectory() as tmp:
            raw = os.path.join(tmp, "raw")
            write_case(raw, case)
            for rep, script in DRIVER_SCRIPTS.items():
                target = os.path.join(tmp, "out_" + rep)
                rh.run_script(script, ["-raw_dir", raw, "-label_dir", raw, "-target_dir", target, "-dataset", flag])
                result["%s/%s" % (tag, rep)] = _digest_tree(target)
    torch.set_num_threads(os.cpu_count())
    with open(os.path.join(GOLDEN, "drivers_digest.json"), "w") as fh:
        json.dump(result, fh, indent=1, sort_keys=True)
    return result


# (img_size, input_img_size, sr, crop fractions in [0,1) of the admissible range, flip)
DATASET_CASES = [((32, 40), (32, 40), 1.0, 0.0, 0.0, False),
                 ((32, 40), (32, 40), 1.0, 0.0, 0.0, True),
                 ((32, 40), (32, 40), 1.3, 0.35, 0.8, False),
                 ((32, 40), (32, 40), 1.4999, 0.999, 0.01, True),
                 ((32, 40), (48, 56), 1.17, 0.5, 0.5, True),
                 ((24, 36), (32, 40), 1.0, 0.0, 0.0, False)]


def dataset_golden():
    """Run the unmodified ``data/dataset.py:propheseeTafDataset.__getitem__`` on a tiny on-disk
    dataset with the random draws of its augmentation pinned, and record inputs and outputs."""
    import random
    import sys
    import types
    import numpy.lib.format as nf
    sys.path.insert(0, rh.REFERENCE_ROOT)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if not hasattr(nf, "_read_array_header"):                       # numpy-2 shim
        nf._read_array_header = lambda fp, version: (nf.read_array_header_1_0(fp) if version == (1, 0)
                                                     else nf.read_array_header_2_0(fp))
    import importlib
    ds_mod = importlib.import_module("data.dataset")
    out = {}
    K = 8
    saved = (random.random, random.uniform)
    try:
        for i, (img_size, in_size, sr, fx, fy, flip) in enumerate(DATASET_CASES):
            with tempfile.TemporaryDirectory() as tmp:
                label_dir, data_dir = os.path.join(tmp, "labels"), os.path.join(tmp, "taf")
                os.makedirs(os.path.join(label_dir, "train"))
                t_label = 100000 + 50000 * i
                # boxes that survive every crop, so that the augmentation loop ends at its first draw
                boxes = np.zeros(2, dtype=synth.BBOX_DTYPE)
                boxes["t"] = t_label
                boxes["x"], boxes["y"], boxes["w"], boxes["h"] = 100.0, 80.0, 120.0, 90.0
                np.save(os.path.join(label_dir, "train", "rec_bbox.npy"), boxes)
                rng = np.random.default_rng(700 + i)
                files = {}
                for sub in ("bins4", "bins8"):
                    os.makedirs(os.path.join(data_dir, "train", sub))
                    files[sub] = rng.integers(0, 256, (K, img_size[0], img_size[1]), dtype=np.uint8)
                    files[sub].tofile(os.path.join(data_dir, "train", sub, "rec_%d.npy" % t_label))
                Hin, Win = in_size
                lo_x, lo_y = int(Win - sr * Win), int(Hin - sr * Hin)
                cx_draw, cy_draw = lo_x * (1.0 - fx), lo_y * (1.0 - fy)      # a point of uniform(lo, 0)
                draws = {"random": [0.1 if sr > 1.0 else 0.9, 0.1 if flip else 0.9], "uniform": [sr, cx_draw, cy_draw], "r": 0, "u": 0}

                def fake_random():
                    v = draws["random"][draws["r"] % 2]
                    draws["r"] += 1
                    return v

                def fake_uniform(a, b):
                    v = draws["uniform"][draws["u"] % 3]
                    draws["u"] += 1
                    assert min(a, b) <= v <= max(a, b), (a, b, v)
                    return v
                random.random, random.uniform = fake_random, fake_uniform
                ds = ds_mod.propheseeTafDataset(label_dir, data_dir, dataset="gen1", input_img_size=list(in_size),
                                                img_size=list(img_size), event_volume_bins=K, mode="train", augment=True)
                img, _labels, _name, _t = ds[0]
                random.random, random.uniform = saved
                assert draws["r"] == 2 and draws["u"] == (3 if sr > 1.0 else 0), draws      # one pass of the loop
                out["ds%d_bins4" % i], out["ds%d_bins8" % i] = files["bins4"], files["bins8"]
                out["ds%d_params" % i] = np.array([img_size[0], img_size[1], Hin, Win, int(cx_draw) if sr > 1.0 else 0,
                                                   int(cy_draw) if sr > 1.0 else 0, int(flip)], dtype=np.int64)
                out["ds%d_sr" % i] = np.array([sr], dtype=np.float64)
                out["ds%d_out" % i] = np.ascontiguousarray(img).astype(np.float32)
    finally:
        random.random, random.uniform = saved
    np.savez_compressed(os.path.join(GOLDEN, "dataset_read.npz"), **out)
    return out


def timesurface_golden():
    """``generate_opticalflow.py:generate_timesurface`` (a numba-jitted function: compiled here as
    plain Python with an identity ``jit``) on two seeded event sets."""
    import ast
    path = os.path.join(rh.REFERENCE_ROOT, "generate_opticalflow.py")
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "generate_timesurface"]
    ns = {"np": np, "jit": lambda *a, **k: (lambda f: f)}
    exec(compile(tree, path, "exec"), ns)  # noqa: S102 - executing the reference is the point
    out = {}
    for tag, (H, W, n, t_lo, t_hi, seed) in {"a": (24, 32, 4000, 1000, 301000, 5), "b": (16, 20, 300, 0, 60000, 6)}.items():
        rng = np.random.default_rng(seed)
        ev = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), np.sort(rng.integers(t_lo, t_hi, n)),
                       rng.integers(0, 2, n)], 1).astype(np.float64)
        v1, v2 = ns["generate_timesurface"](ev, np.zeros((H, W)), np.zeros((H, W)), float(t_hi))
        out["ts_%s_events" % tag], out["ts_%s_shape" % tag] = ev, np.array([H, W])
        out["ts_%s_v1" % tag], out["ts_%s_v2" % tag] = np.asarray(v1), np.asarray(v2)
    np.savez_compressed(os.path.join(GOLDEN, "timesurface.npz"), **out)
    return out


def main():
    assert rh.available(), "reference not mounted at " + rh.REFERENCE_ROOT
    os.makedirs(GOLDEN, exist_ok=True)
    enc = encoder_golden()
    drv = driver_golden()
    dsg = dataset_golden()
    tsg = timesurface_golden()
    print("dataset_read.npz:", len(dsg), "arrays;", "timesurface.npz:", len(tsg), "arrays")
    print("encoders_small.npz:", len(enc), "arrays;", "drivers_digest.json:",
          sum(len(v) for v in drv.values()), "files")


if __name__ == "__main__":
    main()
