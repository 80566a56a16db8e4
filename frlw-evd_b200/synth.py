"""Synthetic Poisson-plus-moving-edge event streams and the
writers that put them on disk in the Prophesee formats the reference reads.

The stream definition is the benchmark contract: ``numpy.random.Generator(PCG64(seed))``,
sorted u32 microsecond timestamps (ties allowed), 20 % uniform background events and
80 % events on four full-height moving edges.  Recording ``r`` uses seed ``1000 + r``.

Writers: ``write_dat`` emits the 8-byte Event2D records decoded by the reference at
``src/io/dat_events_tools.py:82-100`` (with or without the ``% `` text header, see
``parse_header`` :118-173); ``write_bbox_npy`` emits the structured label array whose
header ``src/io/npy_events_tools.py:30-61`` parses.
"""
from __future__ import annotations

import os

import numpy as np

SENSORS = {
    "gen1": (240, 304),
    "gen4": (720, 1280),
    "1mp": (720, 1280),
}

EVENT_DTYPE = np.dtype([("t", "<u4"), ("x", "<u2"), ("y", "<u2"), ("p", "u1")])

# Label record: the 8 fields of evaluate/src/io/box_loading.py:14, PACKED (36 bytes).  The
# reference re-builds the dtype from (name, format) pairs (src/io/npy_events_tools.py:54)
# and therefore only reads label files whose records carry no padding.
BBOX_DTYPE = np.dtype([("t", "<i8"), ("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"),
                       ("class_id", "<u4"), ("track_id", "<u4"), ("class_confidence", "<f4")])


def make_stream(height: int, width: int, duration_us: int, rate_eps: float, seed: int,
                n_edges: int = 4, background: float = 0.2):
    """Return SoA arrays ``(t u32, x u16, y u16, p u8)`` of a synthetic recording."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(round(rate_eps * duration_us / 1e6))
    t = np.sort(rng.integers(0, duration_us, n, dtype=np.uint32))
    x = np.empty(n, dtype=np.uint16)
    y = rng.integers(0, height, n, dtype=np.uint16)
    p = np.empty(n, dtype=np.uint8)

    is_bg = rng.random(n, dtype=np.float32) < background
    n_bg = int(is_bg.sum())
    x[is_bg] = rng.integers(0, width, n_bg, dtype=np.uint16)
    p[is_bg] = rng.integers(0, 2, n_bg, dtype=np.uint8)

    fg = np.flatnonzero(~is_bg)
    edge = rng.integers(0, n_edges, fg.size)
    x0 = rng.uniform(0, width, n_edges)
    speed = rng.uniform(200.0, 2000.0, n_edges) * rng.choice([-1.0, 1.0], n_edges)  # px/s
    slope = rng.uniform(-0.5, 0.5, n_edges)                                          # px/row
    tf = t[fg].astype(np.float64) * 1e-6
    xe = x0[edge] + speed[edge] * tf + slope[edge] * y[fg].astype(np.float64)
    xe += rng.normal(0.0, 0.7, fg.size)
    x[fg] = np.mod(np.rint(xe), width).astype(np.uint16)
    pol = (speed[edge] > 0)
    flip = rng.random(fg.size, dtype=np.float32) < 0.1
    p[fg] = (pol ^ flip).astype(np.uint8)
    return t, x, y, p


def label_times(duration_us: int, first_us: int = 100_000, period_us: int = 50_000) -> np.ndarray:
    """One synthetic label timestamp every 50 ms starting at 100 ms."""
    return np.arange(first_us, duration_us, period_us, dtype=np.int64)


def pack_dat_records(t, x, y, p) -> np.ndarray:
    """Pack SoA events into the ``.dat`` record ``(u4 t, i4 w)``,
    ``w = x | y << 14 | p << 28`` (inverse of ``dat_events_tools.py:96-98``)."""
    rec = np.empty(len(t), dtype=np.dtype([("t", "<u4"), ("w", "<i4")]))
    rec["t"] = t
    w = x.astype(np.uint32) | (y.astype(np.uint32) << 14) | (p.astype(np.uint32) << 28)
    rec["w"] = w.view(np.int32)
    return rec


def write_dat(path: str, t, x, y, p, height=None, width=None, header: bool = True) -> None:
    """Write a ``*_td.dat`` file.  ``header=False`` gives the headerless variant
    produced by ``sampling_dataset.py:59,112-116`` (ev_type 0 / ev_size 8 implied)."""
    with open(path, "wb") as fh:
        if header:
            fh.write(b"% Data file containing Event2D events.\n% Version 2\n")
            fh.write(b"% Date 2026-01-01 00:00:00\n")
            if height is not None:
                fh.write(("%% Height %d\n" % height).encode("latin-1"))
            if width is not None:
                fh.write(("%% Width %d\n" % width).encode("latin-1"))
            fh.write(bytes([0, 8]))  # ev_type, ev_size
        pack_dat_records(t, x, y, p).tofile(fh)


def write_bbox_npy(path: str, times, legacy_names: bool = False) -> None:
    """Write a ``*_bbox.npy`` label file holding one box per timestamp (two for every
    third timestamp, so ``np.unique`` has something to do)."""
    times = np.asarray(times, dtype=np.int64)
    dup = times[::3]
    all_t = np.sort(np.concatenate([times, dup]))
    boxes = np.zeros(all_t.size, dtype=BBOX_DTYPE)
    boxes["t"] = all_t
    boxes["x"], boxes["y"], boxes["w"], boxes["h"] = 10.0, 12.0, 40.0, 30.0
    boxes["class_confidence"] = 1.0
    if legacy_names:
        names = list(boxes.dtype.names)
        names[names.index("t")] = "ts"
        names[names.index("class_confidence")] = "confidence"
        boxes = boxes.view(np.dtype([(n, boxes.dtype[i]) for i, n in enumerate(names)]))
    np.save(path, boxes)


def write_recording(root_raw: str, root_label: str, mode: str, name: str, sensor: str,
                    duration_us: int, rate_eps: float, seed: int, header: bool = True):
    """Materialise one synthetic recording as ``<raw>/<mode>/<name>_td.dat`` and
    ``<label>/<mode>/<name>_bbox.npy``; returns the SoA arrays and label times."""
    h, w = SENSORS[sensor]
    t, x, y, p = make_stream(h, w, duration_us, rate_eps, seed)
    os.makedirs(os.path.join(root_raw, mode), exist_ok=True)
    os.makedirs(os.path.join(root_label, mode), exist_ok=True)
    write_dat(os.path.join(root_raw, mode, name + "_td.dat"), t, x, y, p, h, w, header=header)
    labels = label_times(duration_us)
    write_bbox_npy(os.path.join(root_label, mode, name + "_bbox.npy"), labels)
    return (t, x, y, p), labels
