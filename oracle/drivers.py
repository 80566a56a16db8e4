"""CPU restatement of the four ``generate_*.py`` drivers (windowing + file layout).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the unmodified
reference scripts, run end to end in the build container, by
``tests/test_oracle_vs_reference.py`` and by the file digests in ``tests/golden/``.
Citations are relative to ``/root/reference``.

Each ``run_*`` takes the reference's four CLI arguments and writes the same files:
headerless raw uint8, C order, named ``<file>_<label t>.npy`` (reader contract:
``data/dataset.py:241-249,294-308``).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import encoders as enc
from .psee_io import Loader, read_label_times

GEOMETRY = {"gen4": ((720, 1280), (512, 640))}
DEFAULT_GEOMETRY = ((240, 304), (256, 320))          # every other -dataset value


def _geometry(dataset):
    return GEOMETRY.get(dataset, DEFAULT_GEOMETRY)


def _recordings(raw_dir, label_dir):
    """modes -> files -> (loader, label times); ``generate_taf.py:112-153``."""
    for mode in ("train", "val", "test"):
        try:
            listing = os.listdir(os.path.join(raw_dir, mode))
        except Exception:
            continue
        for name in [f[:-7] for f in listing if f[-3:] == "dat"]:
            labels = read_label_times(os.path.join(label_dir, mode, name + "_bbox.npy"))
            yield mode, name, Loader(os.path.join(raw_dir, mode, name + "_td.dat")), labels


def _stage(ev) -> torch.Tensor:
    """Host staging D4 (``generate_taf.py:195``): float64 ``[N,4]`` columns (x, y, t, p)."""
    arr = np.stack([ev["x"], ev["y"], ev["t"], ev["p"]], axis=1).astype(np.float64)
    return torch.from_numpy(arr)


def _dump(tensor, *path):
    os.makedirs(os.path.join(*path[:-1]), exist_ok=True)
    np.asarray(tensor).astype(np.uint8).tofile(os.path.join(*path))


# --------------------------------------------------------------------------- E2
def run_count_image(raw_dir, label_dir, target_dir, dataset="gen4"):
    """``generate_eventcountimage.py:67-189``."""
    shape, target = _geometry(dataset)
    windows = [400000, 800000, 1200000] if dataset == "gen4" else [50000, 100000, 200000]
    rh, rw = target[0] / shape[0], target[1] / shape[1]
    biggest = max(windows)
    for mode, name, loader, labels in _recordings(raw_dir, label_dir):
        upper, carry = -100000000, None
        for label in labels:
            end_count = loader.seek_time(int(label))
            if end_count is None:
                continue
            start_count = max(int(end_count - biggest), 0)
            if start_count <= upper:
                start_count = upper
            loader.seek_event(start_count)
            events = _stage(loader.load_n_events(int(end_count - start_count)))
            if carry is not None:
                events = torch.cat([carry, events])
            carry = events[-biggest:]
            upper = end_count
            for n in windows:
                ev = events[-n:].clone()
                if target[0] < shape[0]:
                    ev[:, 0] *= rw
                    ev[:, 1] *= rh
                    vol = enc.count_image(ev, target)
                else:
                    vol = enc.nearest_resize(enc.count_image(ev, shape), target)
                _dump(vol.numpy(), target_dir, "EventCountImage{0}".format(n), mode,
                      name + "_" + str(label) + ".npy")


# --------------------------------------------------------------------------- A2
def run_sae(raw_dir, label_dir, target_dir, dataset="gen4"):
    """``generate_surfaceofactiveevents.py:82-220``."""
    shape, target = _geometry(dataset)
    lambdas = [0.00001, 0.0000025, 0.000001]
    sub_windows = [554126, 2216505, 5541263]
    span = 5000000
    rh, rw = target[0] / shape[0], target[1] / shape[1]
    for mode, name, loader, labels in _recordings(raw_dir, label_dir):
        t_upper, c_upper, memory = -100000000, 0, None
        for label in labels:
            end_time = int(label)
            end_count = loader.seek_time(end_time)
            if end_count is None:
                continue
            start_time = end_time - span
            start_count = loader.seek_time(0 if start_time < 0 else start_time)
            if start_count is None or start_time < 0:
                start_count = 0
            if start_time <= t_upper:
                start_count = c_upper
            loader.seek_event(start_count)
            events = _stage(loader.load_n_events(int(end_count - start_count)))
            t_upper, c_upper = label, end_count
            keep = None
            for tw in (sub_windows if mode == "test" else [max(sub_windows)]):
                ev = events[events[:, 2] > end_time - tw].clone()
                if target[0] < shape[0]:
                    ev[:, 0] *= rw
                    ev[:, 1] *= rh
                    vol, memory = enc.sae_surfaces(ev, target, lambdas, memory, label)
                else:
                    vol, memory = enc.sae_surfaces(ev, shape, lambdas, memory, label)
                    vol = enc.nearest_resize(vol, target)
                vol = vol.view(len(lambdas), 2, target[0], target[1])
                if tw == max(sub_windows):
                    keep = vol
            for j, lam in enumerate(lambdas):
                _dump(keep[j].numpy(), target_dir, "SurfaceOfActiveEvents{0}".format(lam), mode,
                      name + "_" + str(label) + ".npy")


# --------------------------------------------------------------------------- V2
def run_event_volume(raw_dir, label_dir, target_dir, dataset="gen1"):
    """``generate_eventvolume.py:58-175``."""
    shape, target = _geometry(dataset)
    windows = [250000, 500000, 1000000]
    K = 5
    rh, rw = target[0] / shape[0], target[1] / shape[1]
    for mode, name, loader, labels in _recordings(raw_dir, label_dir):
        for label in labels:
            end_time = int(label)
            if loader.seek_time(end_time) is None:
                break
            start_time = int(end_time - max(windows))
            if start_time > 0:
                loader.seek_time(start_time)
                raw = loader.load_delta_t(end_time - start_time)
            else:
                loader.seek_time(0)
                raw = loader.load_delta_t(end_time)
            staged = _stage(raw)[-10000000:]
            for tw in windows:
                ev = staged[staged[:, 2] > end_time - tw]
                ev[:, 2] = (ev[:, 2] - (end_time - tw)) / tw
                if target[0] < shape[0]:
                    ev[:, 0] *= rw
                    ev[:, 1] *= rh
                    vol = enc.event_volume(ev, target, K)
                else:
                    vol = enc.nearest_resize(enc.event_volume(ev, shape, K), target)
                vol = vol.numpy()
                _dump(np.where(vol > 255, 255, vol), target_dir, "EventVolume{0}".format(tw), mode,
                      name + "_" + str(label) + ".npy")


# --------------------------------------------------------------------------- T2
def taf_window_plan(loader, label, t_upper, c_upper, abin=10000, span=80000, min_events=50000000):
    """Window of one label timestamp -- ``generate_taf.py:160-187``.  Returns
    ``None`` (skip) or ``(fresh, start_time, end_time, start_count, end_count)``."""
    end_time = int(label)
    end_count = loader.seek_time(end_time)
    if end_count is None:
        return None
    start_count = max(end_count - min_events, 0)
    loader.seek_event(start_count)
    start_time = int(loader.current_time)
    if end_time - start_time < span:
        start_time = end_time - span
    else:
        start_time = end_time - round((end_time - start_time - span) / abin) * abin - span
    if start_time > t_upper:
        start_count = loader.seek_time(start_time)
        if start_count is None or start_time < 0:
            start_count = 0
        return True, start_time, end_time, start_count, end_count
    start_count, start_time = c_upper, t_upper
    end_time = round((end_time - start_time) / abin) * abin + start_time
    if end_time > loader.total_time():
        end_time = loader.total_time()
    end_count = loader.seek_time(end_time)
    return False, start_time, end_time, start_count, end_count


def taf_windows_in_memory(staged, windows, abin, grid, K, scale=None, state=None):
    """The inner loops of ``generate_taf.py:195-222`` on an in-memory staging matrix
    (float64 ``[N,4]`` x, y, t, p) for a list of ``(ev_begin, ev_end, start_time, n_bins,
    fresh)`` windows.  Returns ``(list of [2K,H,W] tensors, final state)``.  Used by the
    parity tests and as the timed CPU baseline of ``bench.py``."""
    outs = []
    if state is None:
        state = enc.taf_fresh_state(grid, K)
    for (lo, hi, start, n_bins, fresh) in windows:
        ev = staged[lo:hi]
        z = torch.zeros_like(ev[:, 0])
        for i in range(n_bins):
            a, b = start + i * abin, start + (i + 1) * abin
            z = torch.where((ev[:, 2] >= a) & (ev[:, 2] <= b), torch.zeros_like(z) + i, z)
        ev = torch.cat([ev, z[:, None]], dim=1)
        if fresh:
            state = enc.taf_fresh_state(grid, K)
        vol = None
        for i in range(n_bins):
            e = ev[ev[:, 4] == i]
            t_min, t_max = start + i * abin, start + (i + 1) * abin
            e[:, 2] = (e[:, 2] - t_min) / (t_max - t_min + 1e-8)
            if scale is not None:
                e[:, 0] *= scale[0]
                e[:, 1] *= scale[1]
            vol, state = enc.taf_bin_update(e, grid, state, K)
        if vol is None:
            vol = state.permute(3, 2, 0, 1).contiguous().view(2 * K, grid[0], grid[1])
        outs.append(vol)
    return outs, state


def run_taf(raw_dir, label_dir, target_dir, dataset="gen4"):
    """``generate_taf.py:78-243``."""
    shape, target = _geometry(dataset)
    abin, K = 10000, 8
    rh, rw = target[0] / shape[0], target[1] / shape[1]
    grid = target if target[0] < shape[0] else shape
    out_root = os.path.join(target_dir, "taf")
    for mode, name, loader, labels in _recordings(raw_dir, label_dir):
        t_upper, c_upper, state, vol = -1e16, -1, None, None
        for label in labels:
            plan = taf_window_plan(loader, label, t_upper, c_upper, abin, abin * K)
            if plan is None:
                continue
            fresh, start_time, end_time, start_count, end_count = plan
            loader.seek_event(start_count)
            events = _stage(loader.load_n_events(int(end_count - start_count)))
            n_bins = math.ceil((end_time - start_time) / abin)
            z = torch.zeros_like(events[:, 0])
            for i in range(n_bins):         # inclusive edges, later bin wins (:201-202)
                lo, hi = start_time + i * abin, start_time + (i + 1) * abin
                z = torch.where((events[:, 2] >= lo) & (events[:, 2] <= hi), torch.zeros_like(z) + i, z)
            events = torch.cat([events, z[:, None]], dim=1)
            if fresh:
                state = enc.taf_fresh_state(grid, K)
            for i in range(n_bins):
                ev = events[events[:, 4] == i]
                t_min, t_max = start_time + i * abin, start_time + (i + 1) * abin
                ev[:, 2] = (ev[:, 2] - t_min) / (t_max - t_min + 1e-8)
                if target[0] < shape[0]:
                    ev[:, 0] *= rw
                    ev[:, 1] *= rh
                    vol, state = enc.taf_bin_update(ev, target, state, K)
                else:
                    vol, state = enc.taf_bin_update(ev, shape, state, K)
                    vol = enc.nearest_resize(vol, target)
            # n_bins == 0 re-uses the previous (already transformed) `vol`: latent bug kept.
            vol = enc.leaky_transform(vol.view(K, 2, target[0], target[1]))
            newest_first = np.flip(vol.numpy().copy(), axis=0)
            fname = name + "_" + str(label) + ".npy"
            _dump(newest_first[:4], out_root, mode, "bins{0}".format(K // 2), fname)
            _dump(newest_first[4:], out_root, mode, "bins{0}".format(K), fname)
            t_upper, c_upper = end_time, end_count
