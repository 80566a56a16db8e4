"""Surface of Active Events -- drop-in for ``generate_surfaceofactiveevents.py``.

``taf_cuda`` (the reference's name for the SAE kernel, :44-69), ``generate_leaky_cuda``
(:71-80) and the Event Volume copy (:18-42) keep their signatures; the command line
reproduces the driver (:82-220).

The reference scatters timestamps with a non-accumulating ``index_put_`` whose result for
duplicate pixels is unspecified on CUDA (and races across CPU threads); this library
defines it as the maximum timestamp, i.e. the sequential result for time-sorted input.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ops
from .recordings import DeviceRecording, Geometry, iter_recordings, parse_args

LAMDAS = [0.00001, 0.0000025, 0.000001]            # :103
TIME_WINDOW = [554126, 2216505, 5541263]           # :104 (timed sub-windows, test mode)
EVENTS_WINDOW = 5000000                            # :106


def generate_agile_event_volume_cuda(events, shape, events_window=50000, volume_bins=5):
    """``generate_surfaceofactiveevents.py:18-42`` (Event Volume copy, returns the tensor only)."""
    return ops.event_volume_aos64(events, tuple(shape), int(volume_bins))


def taf_cuda(x, y, t, p, shape, lamdas, memory, now):
    """``:44-69``: ``x, y, p`` integer tensors, ``t`` float32 timestamps.
    Returns ``(f32 [2L,H,W], f32 memory [2,H,W], seconds)``."""
    events = torch.stack([x.double(), y.double(), t.double(), p.double()], dim=1)
    tick = time.time()
    out, mem = ops.sae_aos64(events, tuple(shape), list(lamdas), memory, now)
    torch.cuda.synchronize()
    return out, mem, time.time() - tick


def generate_leaky_cuda(events, shape, lamdas, memory, now):
    """``:71-80``: float64 ``[N,4]`` (x, y, t, p); events outside the grid are dropped."""
    tick = time.time()
    out, mem = ops.sae_aos64(events, tuple(shape), list(lamdas), memory, now)
    torch.cuda.synchronize()
    return out, mem, time.time() - tick


def encode_recording(rec: DeviceRecording, labels, geom: Geometry, mode="train"):
    """Yield ``(label, u8 [L,2,Ht,Wt])`` (:147-213)."""
    loader = rec.loader
    t_upper, c_upper, memory = -100000000, 0, None
    for label in labels:
        end_time = int(label)
        end_count = loader.seek_time(end_time)
        if end_count is None:
            continue
        start_time = end_time - EVENTS_WINDOW
        start_count = loader.seek_time(0 if start_time < 0 else start_time)
        if start_count is None or start_time < 0:
            start_count = 0
        if start_time <= t_upper:
            start_count = c_upper
        t_upper, c_upper = label, end_count
        u8 = None
        for tw in (TIME_WINDOW if mode == "test" else [max(TIME_WINDOW)]):
            lo = loader.upper_index(end_time - tw, start_count, end_count)     # events[:, 2] > end_time - tw
            u8, memory = ops.sae_u8(rec.events.slice(lo, end_count), geom.grid, geom.target, LAMDAS, memory, label,
                                    geom.coord_maps, geom.resize_maps)         # the largest window comes last
        yield label, u8


def plan_windows(loader, labels):
    """The driver's window of every label (:147-175) as ``(label, ev_begin, ev_end)``; in test
    mode the reference also encodes two shorter sub-windows first, which only time the encoder:
    they are subsets of the largest one and leave the same state behind.  All labels at once
    (numpy): the per-label ``seek_time`` calls of the reference are index look-ups here."""
    labels = np.asarray(labels, dtype=np.int64)
    end_count = loader.seek_index_many(labels)
    keep = end_count >= 0                                    # `if end_count is None: continue`
    labels, end_count = labels[keep], end_count[keep]
    if labels.size == 0:
        return []
    start_time = labels - EVENTS_WINDOW
    start_count = loader.seek_index_many(np.where(start_time < 0, 0, start_time))
    start_count = np.where((start_count < 0) | (start_time < 0), 0, start_count)
    # a window that would reach back before the previous label starts where that one ended (:165-167)
    prev_label = np.concatenate([[-100000000], labels[:-1]])
    prev_end = np.concatenate([[0], end_count[:-1]])
    start_count = np.where(start_time <= prev_label, prev_end, start_count)
    lo = loader.upper_index_many(labels - max(TIME_WINDOW), start_count, np.maximum(end_count, start_count))
    return [(labels[i], int(lo[i]), int(end_count[i])) for i in range(labels.size)]


def encode_chunks(rec: DeviceRecording, labels, geom: Geometry, labels_per_call=256):
    """Whole-recording form of ``encode_recording``: one bucketing + tile-kernel call per ``labels_per_call``
    labels instead of two launches per label.  Yields ``(labels of the chunk, u8 [n_labels, L, 2, Ht, Wt])``.
    Falls back to the per-label path when the windows of consecutive labels overlap (labels less than one
    window apart never do)."""
    plan = plan_windows(rec.loader, labels)
    if any(b[1] < a[2] for a, b in zip(plan, plan[1:])):
        per_label = list(encode_recording(rec, labels, geom))
        for first in range(0, len(per_label), labels_per_call):
            part = per_label[first:first + labels_per_call]
            yield [label for label, _ in part], torch.stack([u8 for _, u8 in part])
        return
    time_of = rec.loader.time_of
    memory = None
    for first in range(0, len(plan), labels_per_call):
        part = plan[first:first + labels_per_call]
        windows = [(lo, hi, label, time_of(lo) if hi > lo else 0, time_of(hi - 1) if hi > lo else 0) for label, lo, hi in part]
        latest, memory = ops.sae_stream(rec.events, windows, geom.grid, memory, geom.coord_maps)
        yield [label for label, _, _ in part], ops.sae_decay_u8_batch(latest, [w[2] for w in windows], LAMDAS, geom.target,
                                                                      geom.resize_maps)


def encode_recording_stream(rec: DeviceRecording, labels, geom: Geometry, labels_per_call=256):
    """``encode_chunks`` label by label: yields the same ``(label, u8 [L,2,Ht,Wt])`` as ``encode_recording``."""
    for chunk_labels, u8 in encode_chunks(rec, labels, geom, labels_per_call):
        for i, label in enumerate(chunk_labels):
            yield label, u8[i]


def main(argv=None):
    from .recordings import AsyncWriter, PinnedRing
    args = parse_args("gen4", argv)
    geom = Geometry.for_dataset(args.dataset)
    writer, ring = AsyncWriter(), PinnedRing()
    total_time, total_count = 0.0, 0
    for mode, name, event_file, labels in iter_recordings(args.raw_dir, args.label_dir):
        rec = DeviceRecording(event_file)
        torch.cuda.synchronize()
        tick, count = time.time(), 0
        for chunk_labels, u8 in encode_chunks(rec, labels, geom):
            def emit(host, names=[name + "_" + str(label) + ".npy" for label in chunk_labels], mode=mode):
                return [writer.put(host[i, j], args.target_dir, "SurfaceOfActiveEvents{0}".format(lam), mode, fname)
                        for i, fname in enumerate(names) for j, lam in enumerate(LAMDAS)]
            ring.push(u8.contiguous(), emit)     # device -> pinned ring -> files, behind the next chunk's kernels
            count += len(chunk_labels)
        if mode == "test":               # the reference times and counts the test split only (:121-123,186-188)
            torch.cuda.synchronize()
            total_time += time.time() - tick
            total_count += count
    ring.flush()
    writer.close()
    if total_count and total_time:
        print("Average Representation time: ", total_time / total_count)


if __name__ == "__main__":
    main()
