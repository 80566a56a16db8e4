"""CPU restatement of the reference's Prophesee file I/O (numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the unmodified
reference loader by ``tests/test_oracle_vs_reference.py`` (build container) and by the
driver goldens under ``tests/golden/``.  Citations are relative to ``/root/reference``.

The restatement is index based (the whole record array is memory-mapped) but keeps
the reference's cursor semantics, including its quirks: ``seek_time`` may return a
non-leftmost index on an exact probe hit and then leaves the cursor one event late.
"""
from __future__ import annotations

import bisect

import numpy as np

RECORD = np.dtype([("t", "<u4"), ("w", "<i4")])                       # dat_events_tools.py:16
DECODED = np.dtype([("t", "<u4"), ("x", "<u2"), ("y", "<u2"), ("p", "u1")])  # psee_loader.py:39-44


def parse_dat_header(path):
    """``dat_events_tools.py:118-173``: lines starting ``"% "`` are the header; if there
    was at least one, two bytes ``ev_type, ev_size`` follow, else the file is headerless
    with ev_type 0 / ev_size 8.  Returns ``(data_offset, ev_type, ev_size, (H, W))``."""
    size = [None, None]
    n_lines = 0
    with open(path, "rb") as fh:
        while True:
            pos = fh.tell()
            line = fh.readline()
            if line.decode("latin-1")[:2] != "% ":
                break
            words = line.split()
            if len(words) > 1:
                if words[1] == b"Height":
                    size[0] = int(words[2])
                if words[1] == b"Width":
                    size[1] = int(words[2])
            n_lines += 1
        fh.seek(pos)
        if n_lines > 0:
            ev_type, ev_size = fh.read(1)[0], fh.read(1)[0]
        else:
            ev_type, ev_size = 0, 8
        return fh.tell(), int(ev_type), int(ev_size), tuple(size)


def decode_records(rec: np.ndarray) -> np.ndarray:
    """``dat_events_tools.py:92-100``: ``x = w & 0x3FFF``, ``y = (w & 0x0FFFC000) >> 14``,
    ``p = (w & 0x10000000) >> 28`` on the signed 32-bit word."""
    out = np.empty(rec.shape[0], dtype=DECODED)
    out["t"] = rec["t"]
    w = rec["w"]
    out["x"] = np.bitwise_and(w, 16383)
    out["y"] = np.right_shift(np.bitwise_and(w, 268419072), 14)
    out["p"] = np.right_shift(np.bitwise_and(w, 268435456), 28)
    return out


def read_label_times(path) -> np.ndarray:
    """``npy_events_tools.py:30-61`` + ``np.unique(dat_bbox['t'])``
    (``generate_taf.py:146-151``).  Like the reference, the record dtype is re-built
    PACKED from the header's (name, format) pairs (:54) with ``ts -> t`` and
    ``confidence -> class_confidence`` (:56-57), and the payload is read with
    ``np.fromfile`` -- so padded record layouts are mis-read exactly as there."""
    fmt = np.lib.format
    with open(path, "rb") as fh:
        version = fmt.read_magic(fh)
        reader = fmt.read_array_header_1_0 if tuple(version) == (1, 0) else fmt.read_array_header_2_0
        _shape, fortran, dtype = reader(fh)
        assert not fortran, "Fortran order arrays not supported"
        rename = {"ts": "t", "confidence": "class_confidence"}
        fields = [(rename.get(n, n), str(dtype.fields[n][0])) for n in dtype.names]
        boxes = np.fromfile(fh, dtype=fields, count=-1)
    return np.unique(boxes["t"])


class Loader:
    """Restatement of ``PSEELoader`` for ``.dat`` files (``psee_loader.py:13-252``)."""

    def __init__(self, path):
        assert path.split(".")[-1] == "dat", path
        self._start, self.ev_type, self._ev_size, self._size = parse_dat_header(path)
        assert self._ev_size != 0
        self._rec = np.memmap(path, dtype=RECORD, mode="r", offset=self._start)
        self._ev_count = self._rec.shape[0]
        self._cursor = 0                      # event index of the file cursor
        self.done = False
        self.current_time = 0
        self.duration_s = self.total_time() * 1e-6

    def event_count(self):
        return self._ev_count

    def reset(self):                          # :57-61
        self._cursor, self.done, self.current_time = 0, False, 0

    def total_time(self):                     # :230-249
        return int(self._rec["t"][-1]) if self._ev_count else 0

    def seek_event(self, ev_count):           # :161-183
        ev_count = int(ev_count)
        if ev_count <= 0:
            self._cursor, self.current_time = 0, 0
        elif ev_count >= self._ev_count:
            self._cursor = self._ev_count
            self.current_time = int(self._rec["t"][-1]) + 1
        else:
            self._cursor = ev_count
            self.current_time = int(self._rec["t"][ev_count])
        self.done = self._cursor >= self._ev_count

    def seek_time(self, final_time, term_criterion=100000):   # :185-228
        if final_time > self.total_time():
            self._cursor, self.done = self._ev_count, True
            self.current_time = self.total_time() + 1
            return None
        if final_time <= 0:
            self.reset()
            return None
        low, high = 0, self._ev_count
        while high - low > term_criterion:
            middle = (low + high) // 2
            mid = int(self._rec["t"][middle])   # seek_event(middle) then a 1-record read
            if mid > final_time:
                high = middle
            elif mid < final_time:
                low = middle + 1
            else:
                self._cursor = middle + 1       # the probe read advanced the cursor
                self.current_time = final_time
                self.done = self._cursor >= self._ev_count
                return middle
        idx = bisect.bisect_left(self._rec["t"], final_time, low, high)
        self.seek_event(idx)
        self.current_time = final_time
        self.done = self._cursor >= self._ev_count
        return idx

    def load_n_events(self, ev_count):        # :92-115
        ev_count = int(ev_count)
        pos = self._cursor
        left = self._ev_count - pos
        if ev_count >= left:
            self.done = True
            ev_count = left
            out = decode_records(self._rec[pos:pos + ev_count])
            if ev_count > 0:
                self.current_time = int(out["t"][ev_count - 1]) + 1
            self._cursor = self._ev_count
        else:
            out = decode_records(self._rec[pos:pos + ev_count])
            self.current_time = int(self._rec["t"][pos + ev_count])
            self._cursor = pos + ev_count
        return out

    def load_delta_t(self, delta_t):          # :117-159
        if delta_t < 1:
            raise ValueError("load_delta_t(): delta_t must be at least 1 micro-second: {}".format(delta_t))
        if self.done or self._cursor >= self._ev_count:
            self.done = True
            return np.empty((0,), dtype=DECODED)
        final_time = self.current_time + delta_t
        start = self._cursor
        stop = bisect.bisect_left(self._rec["t"], final_time, start, self._ev_count)
        # the reference reads 100k-event batches until one ends at/after final_time (or
        # EOF) and cuts the LAST batch with searchsorted; for time-sorted files the result
        # is the slice [start, first index with t >= final_time).
        n_batches = (stop - start) // 100000 + 1
        last_read = min(self._ev_count, start + n_batches * 100000)
        tmp_time = int(self._rec["t"][last_read - 1])
        self.current_time = final_time if tmp_time >= final_time else tmp_time + 1
        self._cursor = stop
        self.done = self._cursor >= self._ev_count
        return decode_records(self._rec[start:stop])
