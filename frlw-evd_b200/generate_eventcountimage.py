"""Event Count Image -- drop-in for the reference's ``generate_eventcountimage.py``.

``generate_eventframe`` / ``generate_frame`` keep the reference signatures; the command
line (``python -m frlw_evd_b200.generate_eventcountimage -raw_dir ... -dataset gen4``)
reproduces its driver (:67-189) with the recording resident on the GPU.

Driver note: the reference carries the last max-N events in ``memory`` and encodes
``events[-N:]`` (:148-156); with the whole recording on the device that slice is simply
the index range ``[max(end_count - N, 0), end_count)``, and the three nested windows are
accumulated incrementally (every event is counted once per label).
"""
from __future__ import annotations

import time

import torch

from . import ops
from .recordings import DeviceRecording, Geometry, dump_u8, iter_recordings, parse_args


def generate_eventframe(events, shape):
    """``generate_eventcountimage.py:19-41``: ``events`` float64 ``[N,4]`` (x, y, t, p) on
    the GPU.  Returns ``(f32 [2,H,W], seconds)``; value = f32 running sum of 0.05 per event,
    clamped to 1, times 255."""
    tick = time.time()
    out = ops.count_image_aos64(events, tuple(shape))
    torch.cuda.synchronize()
    return out, time.time() - tick


def generate_frame(events, shape, events_window=50000, volume_bins=5):
    """``generate_eventcountimage.py:43-65`` (an untimed copy of the Event Volume encoder)."""
    return ops.event_volume_aos64(events, tuple(shape), int(volume_bins))


def windows_for(dataset):
    return [400000, 800000, 1200000] if dataset == "gen4" else [50000, 100000, 200000]   # :81-88


def encode_recording(rec: DeviceRecording, labels, geom: Geometry, sizes):
    """Yield ``(label, [u8 [2,Ht,Wt] per N])`` (:130-182)."""
    for label in labels:
        end_count = rec.loader.seek_time(int(label))
        if end_count is None:
            continue
        lo = max(end_count - max(sizes), 0)
        u8 = ops.count_images_u8(rec.events.slice(lo, end_count), sizes, geom.grid, geom.target, geom.coord_maps,
                                 geom.resize_maps)
        yield label, list(u8)


def main(argv=None):
    args = parse_args("gen4", argv)
    geom = Geometry.for_dataset(args.dataset)
    sizes = windows_for(args.dataset)
    total_time, total_count = 0.0, 0
    for mode, name, event_file, labels in iter_recordings(args.raw_dir, args.label_dir):
        rec = DeviceRecording(event_file)
        torch.cuda.synchronize()
        tick = time.time()
        for label, frames in encode_recording(rec, labels, geom, sizes):
            for n, u8 in zip(sizes, frames):
                dump_u8(u8, args.target_dir, "EventCountImage{0}".format(n), mode, name + "_" + str(label) + ".npy")
            total_count += 1
        if mode == "test":
            total_time += time.time() - tick
    if total_count and total_time:
        print("Average Representation time: ", total_time / total_count)


if __name__ == "__main__":
    main()
