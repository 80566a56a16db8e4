"""``.dat`` (Event2D) files -- same API as the reference's ``src/io/dat_events_tools.py``.

Format (reference :16,:118-173,:92-100): zero or more text lines starting with ``"% "``
(``Height`` / ``Width`` are parsed), then -- only when there was at least one such line --
two bytes ``ev_type, ev_size``; headerless files imply type 0 / size 8.  Records are
little-endian ``(u4 t, i4 w)`` with ``x = w & 0x3FFF``, ``y = (w >> 14) & 0x3FFF``,
``p = (w >> 28) & 1``.

Decoding of whole recordings happens on the GPU (``ops.decode_dat``); the numpy decode
here serves the host-side ``PSEELoader`` API and small reads.
"""
from __future__ import annotations

import datetime
import os

import numpy as np

EV_TYPE = [("t", "u4"), ("_", "i4")]  # Event2D
EV_STRING = "Event2D"
RECORD_DTYPE = np.dtype([("t", "<u4"), ("_", "<i4")])
DECODED_DTYPE = np.dtype([("t", "<u4"), ("x", "<u2"), ("y", "<u2"), ("p", "u1")])

X_MASK, Y_MASK, P_MASK = 0x3FFF, 0x0FFFC000, 0x10000000


def parse_header(f):
    """Return ``(data offset, ev_type, ev_size, (height, width))`` of an open ``.dat`` file."""
    f.seek(0, os.SEEK_SET)
    size = [None, None]
    comment_lines = 0
    while True:
        pos = f.tell()
        line = f.readline()
        if line[:2] != b"% ":
            break
        words = line.split()
        if len(words) > 2 and words[1] == b"Height":
            size[0] = int(words[2])
        if len(words) > 2 and words[1] == b"Width":
            size[1] = int(words[2])
        comment_lines += 1
    f.seek(pos, os.SEEK_SET)
    if comment_lines > 0:
        ev_type, ev_size = f.read(1)[0], f.read(1)[0]
    else:
        ev_type, ev_size = 0, int(RECORD_DTYPE.itemsize)
    return f.tell(), int(ev_type), int(ev_size), size


def unpack(records: np.ndarray, out: np.ndarray = None) -> np.ndarray:
    """Decode packed records into the ``(t, x, y, p)`` structured layout."""
    n = records.shape[0]
    if out is None:
        out = np.empty(n, dtype=DECODED_DTYPE)
    w = records["_"].view(np.uint32) if records["_"].dtype != np.uint32 else records["_"]
    out["t"][:n] = records["t"]
    out["x"][:n] = w & X_MASK
    out["y"][:n] = (w & Y_MASK) >> 14
    out["p"][:n] = (w & P_MASK) >> 28
    return out


def stream_td_data(file_handle, buffer, dtype, ev_count=-1):
    """Read ``ev_count`` records from the open file into the pre-allocated ``buffer``."""
    dat = np.fromfile(file_handle, dtype=RECORD_DTYPE if dtype == EV_TYPE else dtype, count=ev_count)
    if "_" in dat.dtype.names:
        unpack(dat, buffer)
    else:
        for name in dat.dtype.names:
            buffer[name][:len(dat)] = dat[name]


def load_td_data(filename, ev_count=-1, ev_start=0):
    """Load events ``[ev_start, ev_start + ev_count)`` of a file as a structured array
    with fields ``t, x, y, p`` (the reference returns int16 x / y / p)."""
    with open(filename, "rb") as f:
        _, _, ev_size, _ = parse_header(f)
        if ev_start > 0:
            f.seek(ev_start * ev_size, 1)
        dat = np.fromfile(f, dtype=RECORD_DTYPE, count=ev_count)
    dec = unpack(dat)
    out = np.empty(dat.shape[0], dtype=[("t", "u4"), ("x", "i2"), ("y", "i2"), ("p", "i2")])
    for name in ("t", "x", "y", "p"):
        out[name] = dec[name]
    return out


def count_events(filename):
    with open(filename, "rb") as f:
        bod, _, ev_size, _ = parse_header(f)
        f.seek(0, os.SEEK_END)
        eod = f.tell()
        if (eod - bod) % ev_size != 0:
            raise Exception("unexpected format !")
        return (eod - bod) // ev_size


def write_header(filename, height=240, width=320, ev_type=0):
    """Create a ``.dat`` file with a text header and return the open (binary) handle."""
    if max(height, width) > 2 ** 14 - 1:
        raise ValueError("Coordinates value exceed maximum range in binary .dat file format "
                         "max({:d},{:d}) vs 2^14 - 1".format(height, width))
    f = open(filename, "wb")
    now = datetime.datetime.now(datetime.timezone.utc)
    f.write(("% Data file containing {:s} events.\n% Version 2\n".format(EV_STRING)).encode("latin-1"))
    f.write(now.strftime("%% Date %Y-%m-%d %H:%M:%S\n").encode("latin-1"))
    f.write("% Height {:d}\n% Width {:d}\n".format(height, width).encode("latin-1"))
    f.write(bytes([ev_type, RECORD_DTYPE.itemsize]))
    f.flush()
    return f


def write_event_buffer(f, buffers):
    """Append events (fields ``t, x, y, p``) to an open ``.dat`` file."""
    rec = np.empty(len(buffers["t"]), dtype=RECORD_DTYPE)
    rec["t"] = buffers["t"]
    w = (buffers["x"].astype(np.uint32) | (buffers["y"].astype(np.uint32) << 14) |
         ((buffers["p"] == 1).astype(np.uint32) << 28))
    rec["_"] = w.view(np.int32)
    rec.tofile(f)
    f.flush()
