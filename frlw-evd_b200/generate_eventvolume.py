"""Event Volume -- drop-in for the reference's ``generate_eventvolume.py``.

``generate_agile_event_volume_cuda`` (:15-42) and ``denseToSparse`` (:44-56) keep their
signatures; the command line reproduces the driver (:58-175): for every label the last
second of events is selected with ``seek_time`` + ``load_delta_t`` semantics (index ranges
only -- nothing is decoded on the host), capped to the last 10 M events, and encoded for
the 250 / 500 / 1000 ms windows with K = 5.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ops
from .recordings import DeviceRecording, Geometry, dump_u8, iter_recordings, parse_args

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


def encode_recording(rec: DeviceRecording, labels, geom: Geometry):
    """Yield ``(label, [u8 [2K,Ht,Wt] per window])`` (:118-169); stops at the first label
    past the end of the recording like the reference's ``break``."""
    loader = rec.loader
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
        outs = []
        for tw in TIME_WINDOWS:
            first = loader.upper_index(end_time - tw, lo, hi)         # events_[:, 2] > end_time - tw
            vol = ops.event_volume(rec.events.slice(first, hi), end_time - tw, tw, geom.grid, VOLUME_BINS,
                                   geom.coord_maps)
            outs.append(ops.quantize_u8(geom.to_target(vol), clamp255=True))
        yield label, outs


def main(argv=None):
    args = parse_args("gen1", argv)
    geom = Geometry.for_dataset(args.dataset)
    total_time, total_count = 0.0, 0
    for mode, name, event_file, labels in iter_recordings(args.raw_dir, args.label_dir):
        rec = DeviceRecording(event_file)
        torch.cuda.synchronize()
        tick = time.time()
        for label, outs in encode_recording(rec, labels, geom):
            for tw, u8 in zip(TIME_WINDOWS, outs):
                dump_u8(u8, args.target_dir, "EventVolume{0}".format(tw), mode, name + "_" + str(label) + ".npy")
            total_count += 1
        if mode == "test":
            total_time += time.time() - tick
    if total_count and total_time:
        print("Average Representation time: ", total_time / total_count)


if __name__ == "__main__":
    main()
