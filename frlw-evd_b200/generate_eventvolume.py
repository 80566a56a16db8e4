"""Event Volume -- drop-in for the reference's ``generate_eventvolume.py``.

``generate_agile_event_volume_cuda`` (:15-42) and ``denseToSparse`` (:44-56) keep their
signatures; the command line reproduces the driver (:58-175): for every label the last
second of events is selected with ``seek_time`` + ``load_delta_t`` semantics (index ranges
only -- nothing is decoded on the host), capped to the last 10 M events, and encoded for
the 250 / 500 / 1000 ms windows with K = 5.  The windows nest inside a label and overlap
between labels: the events of a batch of labels are sorted once by sensor tile and every
window splats the records of its own time range (``ops.event_volume_spans``); the uint8
files leave through a ring of pinned buffers and an asynchronous writer.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ops
from .recordings import DeviceRecording, Geometry, iter_recordings, parse_args

TIME_WINDOWS = [250000, 500000, 1000000]     # :82
VOLUME_BINS = 5                              # :83


def generate_agile_event_volume_cuda(events, shape, events_window=50000, volume_bins=5):
    """``:15-42``: float64 ``[N,4]`` (x, y, t_norm, p).  Returns ``(f32 [2K,H,W], seconds)``;
    channel of an event = ``2k + (1 - p)``, scale ``/ 5 * 255`` whatever K is."""
    tick = time.time()
    out = ops.event_volume_aos64(events, tuple(shape), int(volume_bins))
    torch.cuda.synchronize()
    return out, time.time() - tick


def denseToSparse(dense_tensor):
    """``:44-56`` (numpy): indices of the non-zero entries of a 3-D array and their values."""
    nz = np.nonzero(dense_tensor)
    return np.stack(nz), dense_tensor[nz[0], nz[1], nz[2]]


def label_windows(loader, labels):
    """Index ranges of the driver's windows (:118-147): for every label ``[(first, hi, t0, tw) x 3]``.
    Stops at the first label past the end of the recording like the reference's ``break``."""
    plan = []
    for label in labels:
        end_time = int(label)
        if loader.seek_time(end_time) is None:
            break
        start_time = int(end_time - max(TIME_WINDOWS))
        if start_time > 0:
            loader.seek_time(start_time)
            lo, hi = loader.index_delta_t(end_time - start_time)
        else:
            loader.seek_time(0)
            lo, hi = loader.index_delta_t(end_time)
        lo = max(lo, hi - 10000000)                                   # events_[-10000000:]
        # events_[:, 2] > end_time - tw
        plan.append((label, [(loader.upper_index(end_time - tw, lo, hi), hi, end_time - tw, tw) for tw in TIME_WINDOWS]))
    return plan


def encode_windows(rec: DeviceRecording, windows, geom: Geometry, volume_bins=VOLUME_BINS) -> torch.Tensor:
    """u8 ``[n_windows, 2K, Ht, Wt]`` (device) of a list of ``(first, hi, t0, tw)`` windows that may nest and
    overlap: one sort of the events they cover, one splat per window (``ops.event_volume_spans``)."""
    loader = rec.loader
    segments, spans = ops.plan_ev_spans(windows, loader.time_of, loader.lower_index)
    H, W = geom.grid
    if geom.downscale:                 # the kernel writes the file bytes itself
        u8 = torch.empty((len(windows), 2 * volume_bins, H, W), dtype=torch.uint8, device=rec.events.device)
        ops.event_volume_spans(rec.events, segments, spans, geom.grid, volume_bins, geom.coord_maps, out_u8=u8, want_f32=False)
        return u8
    vol = ops.event_volume_spans(rec.events, segments, spans, geom.grid, volume_bins, geom.coord_maps)
    return ops.event_volume_u8_batch(vol, geom.target, geom.resize_maps)


def encode_recording(rec: DeviceRecording, labels, geom: Geometry, labels_per_call=32):
    """Yield ``(label, u8 [3, 2K, Ht, Wt])`` (:118-169), ``labels_per_call`` labels per kernel call."""
    plan = label_windows(rec.loader, labels)
    if plan and not rec.events.is_ordered():
        raise ValueError("the event file is not ordered in time: the loader's seek_time (bisection) needs it")
    for lo in range(0, len(plan), labels_per_call):
        chunk = plan[lo:lo + labels_per_call]
        u8 = encode_windows(rec, [w for _, ws in chunk for w in ws], geom)
        for i, (label, _) in enumerate(chunk):
            yield label, u8[3 * i:3 * i + 3]


def encode_recording_to_files(rec: DeviceRecording, labels, name: str, mode: str, target_dir: str, geom: Geometry, ring,
                              writer, labels_per_call=32) -> int:
    """File layout of :162-167 (``EventVolume<tw>/<mode>/<recording>_<label>.npy``) through the pinned ring and
    the asynchronous writer.  Returns the number of labels encoded."""
    plan = label_windows(rec.loader, labels)
    if plan and not rec.events.is_ordered():
        raise ValueError("the event file is not ordered in time: the loader's seek_time (bisection) needs it")
    for lo in range(0, len(plan), labels_per_call):
        chunk = plan[lo:lo + labels_per_call]
        u8 = encode_windows(rec, [w for _, ws in chunk for w in ws], geom)
        names = [name + "_" + str(label) + ".npy" for label, _ in chunk]

        def emit(host, names=names):
            return [writer.put(host[3 * i + j], target_dir, "EventVolume{0}".format(tw), mode, fname)
                    for i, fname in enumerate(names) for j, tw in enumerate(TIME_WINDOWS)]
        ring.push(u8, emit)
    return len(plan)


def main(argv=None):
    from .recordings import AsyncWriter, PinnedRing
    args = parse_args("gen1", argv)
    geom = Geometry.for_dataset(args.dataset)
    writer, ring = AsyncWriter(), PinnedRing()
    total_time, total_count = 0.0, 0
    for mode, name, event_file, labels in iter_recordings(args.raw_dir, args.label_dir):
        rec = DeviceRecording(event_file)
        torch.cuda.synchronize()
        tick = time.time()
        n = encode_recording_to_files(rec, labels, name, mode, args.target_dir, geom, ring, writer)
        if mode == "test":                 # the reference times and counts the test split only (:104-106,153-155)
            torch.cuda.synchronize()
            total_time += time.time() - tick
            total_count += n
    ring.flush()
    writer.close()
    if total_count and total_time:
        print("Average Representation time: ", total_time / total_count)


if __name__ == "__main__":
    main()
