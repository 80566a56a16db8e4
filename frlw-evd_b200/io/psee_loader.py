"""``PSEELoader`` with the reference's API (``src/io/psee_loader.py:13-252``) on top of a
memory map, plus index-level primitives for the GPU drivers.

The reference moves a file cursor and reads small pieces; here the record array is
memory-mapped once and the cursor is an event index.  Results are identical, including
two behaviours the window drivers depend on:

* ``seek_time`` bisects with one-record probes while the bracket is wider than
  ``term_criterion`` events and returns the probe index on an exact hit -- not
  necessarily the first event with that timestamp -- leaving the cursor one event
  further (:206-219);
* ``load_delta_t`` ends at the first event at/after ``current_time + delta_t`` (:117-159).
"""
from __future__ import annotations

import bisect

import numpy as np

from . import dat_events_tools as dat
from . import npy_events_tools as npy_format


class PSEELoader(object):
    def __init__(self, datfile):
        self._extension = datfile.split(".")[-1]
        assert self._extension in ["dat", "npy"], "input file path = {}".format(datfile)
        with open(datfile, "rb") as fh:
            if self._extension == "dat":
                self._start, self.ev_type, self._ev_size, self._size = dat.parse_header(fh)
                self._dtype = dat.EV_TYPE
                rec_dtype = dat.RECORD_DTYPE
                self._decode_dtype = [("t", "u4"), ("x", "u2"), ("y", "u2"), ("p", "u1")]
            else:
                self._start, self.ev_type, self._ev_size, self._size, _ = npy_format.parse_header(fh)
                self._dtype = self.ev_type
                rec_dtype = np.dtype(self.ev_type)
                self._decode_dtype = list(self.ev_type)
        assert self._ev_size != 0
        self.path = datfile
        self.records = np.memmap(datfile, dtype=rec_dtype, mode="r", offset=self._start)
        self._t = self.records["t"]
        self._ev_count = int(self.records.shape[0])
        self._pos = 0
        self.done = False
        self.current_time = 0
        self.duration_s = self.total_time() * 1e-6

    @classmethod
    def from_records(cls, records: np.ndarray) -> "PSEELoader":
        """A loader over an in-memory ``.dat`` record array (no file behind it)."""
        self = cls.__new__(cls)
        self._extension, self._start, self.ev_type, self._ev_size, self._size = "dat", 0, 0, 8, [None, None]
        self._dtype = dat.EV_TYPE
        self._decode_dtype = [("t", "u4"), ("x", "u2"), ("y", "u2"), ("p", "u1")]
        self.path = None
        self.records = records.view(dat.RECORD_DTYPE) if records.dtype != dat.RECORD_DTYPE else records
        self._t = self.records["t"]
        self._ev_count = int(self.records.shape[0])
        self._pos, self.done, self.current_time = 0, False, 0
        self.duration_s = self.total_time() * 1e-6
        return self

    # -- index-level primitives (used by the GPU drivers) -------------------------------
    def time_of(self, index: int) -> int:
        return int(self._t[index])

    def raw_bytes(self, lo: int = 0, hi: int = None) -> np.ndarray:
        """Payload bytes of events ``[lo, hi)`` (a uint8 view of the memory map)."""
        hi = self._ev_count if hi is None else hi
        return self.records[lo:hi].view(np.uint8)

    @property
    def position(self) -> int:
        """Event index of the cursor."""
        return self._pos

    # -- reference API ---------------------------------------------------------------------
    def reset(self):
        self._pos, self.done, self.current_time = 0, False, 0

    def event_count(self):
        return self._ev_count

    def get_size(self):
        return self._size

    def __repr__(self):
        kind = dat.EV_STRING if self._extension == "dat" else "numpy array element"
        return ("PSEELoader:\n-----------\nEvent Type: {}\nEvent Size: {} bytes\nEvent Count: {}\n"
                "Duration: {} s \n-----------\n").format(kind, self._ev_size, self._ev_count, self.duration_s)

    def _decode(self, lo, hi):
        rec = self.records[lo:hi]
        if self._extension == "dat":
            return dat.unpack(rec, np.empty(hi - lo, dtype=self._decode_dtype))
        return np.array(rec, dtype=self._decode_dtype)

    def load_n_events(self, ev_count):
        ev_count = int(ev_count)
        left = self._ev_count - self._pos
        lo = self._pos
        if ev_count >= left:
            self.done = True
            ev_count = left
            if ev_count > 0:
                self.current_time = int(self._t[lo + ev_count - 1]) + 1
        else:
            self.current_time = int(self._t[lo + ev_count])
        self._pos = lo + ev_count
        return self._decode(lo, lo + ev_count)

    def load_delta_t(self, delta_t):
        lo, hi = self.index_delta_t(delta_t)
        return self._decode(lo, hi)

    def index_delta_t(self, delta_t):
        """``load_delta_t`` without the decode: advances the cursor / clock identically and
        returns the event index range ``[lo, hi)`` it would have loaded."""
        if delta_t < 1:
            raise ValueError("load_delta_t(): delta_t must be at least 1 micro-second: {}".format(delta_t))
        if self.done or self._pos >= self._ev_count:
            self.done = True
            return self._pos, self._pos
        final_time = self.current_time + delta_t
        lo = self._pos
        hi = bisect.bisect_left(self._t, final_time, lo, self._ev_count)   # O(log n) probes of the map
        # timestamp of the last event the reference's 100k-event batching would have read
        batches = (hi - lo) // 100000 + 1
        last_seen = int(self._t[min(self._ev_count, lo + batches * 100000) - 1])
        self.current_time = final_time if last_seen >= final_time else last_seen + 1
        self._pos = hi
        self.done = self._pos >= self._ev_count
        return lo, hi

    # -- the same queries for many timestamps at once (numpy; the cursor does not move) ----------------
    def _bisect_many(self, keys, lo, hi, right):
        """Vectorised ``bisect_left`` / ``bisect_right`` of ``keys`` inside the index ranges ``[lo, hi)``."""
        keys = np.asarray(keys, dtype=np.int64)
        lo = np.broadcast_to(np.asarray(lo, dtype=np.int64), keys.shape).copy()
        hi = np.broadcast_to(np.asarray(hi, dtype=np.int64), keys.shape).copy()
        open_ = lo < hi
        while open_.any():
            mid = (lo + hi) >> 1
            probe = self._t[np.where(open_, mid, 0)].astype(np.int64)
            go_right = (probe <= keys) if right else (probe < keys)
            lo = np.where(open_ & go_right, mid + 1, lo)
            hi = np.where(open_ & ~go_right, mid, hi)
            open_ = lo < hi
        return lo

    def lower_index_many(self, times, lo, hi):
        return self._bisect_many(times, lo, hi, right=False)

    def upper_index_many(self, times, lo, hi):
        return self._bisect_many(times, lo, hi, right=True)

    def seek_index_many(self, times, term_criterion=100000):
        """What ``seek_time`` returns for every entry of ``times`` (``-1`` where it returns ``None``), including
        its coarse phase that stops at the first probe that hits the timestamp exactly (:204-217), which on files
        with repeated timestamps need not be the leftmost one."""
        times = np.asarray(times, dtype=np.int64)
        out = np.full(times.shape, -1, dtype=np.int64)
        live = (times <= self.total_time()) & (times > 0) if self._ev_count else np.zeros(times.shape, dtype=bool)
        lo = np.zeros(times.shape, dtype=np.int64)
        hi = np.full(times.shape, self._ev_count, dtype=np.int64)
        coarse = live & (hi - lo > term_criterion)
        while coarse.any():
            mid = (lo + hi) >> 1
            probe = self._t[np.where(coarse, mid, 0)].astype(np.int64)
            above, below = coarse & (probe > times), coarse & (probe < times)
            hit = coarse & ~above & ~below
            out[hit] = mid[hit]
            live &= ~hit
            hi = np.where(above, mid, hi)
            lo = np.where(below, mid + 1, lo)
            coarse = live & (hi - lo > term_criterion)
        fine = self._bisect_many(times, np.where(live, lo, 0), np.where(live, hi, 0), right=False)
        out[live] = fine[live]
        return out

    def lower_index(self, time_us, lo, hi):
        """First index in ``[lo, hi)`` whose timestamp is ``>= time_us``."""
        return bisect.bisect_left(self._t, time_us, lo, hi)

    def upper_index(self, time_us, lo, hi):
        """First index in ``[lo, hi)`` whose timestamp is ``> time_us`` (the drivers'
        ``events[:, 2] > bound`` filters on time-sorted slices)."""
        return bisect.bisect_right(self._t, time_us, lo, hi)

    def seek_event(self, ev_count):
        ev_count = int(ev_count)
        if ev_count <= 0:
            self._pos, self.current_time = 0, 0
        elif ev_count >= self._ev_count:
            self._pos = self._ev_count
            self.current_time = int(self._t[-1]) + 1
        else:
            self._pos = ev_count
            self.current_time = int(self._t[ev_count])
        self.done = self._pos >= self._ev_count

    def seek_time(self, final_time, term_criterion=100000):
        if final_time > self.total_time():
            self._pos, self.done = self._ev_count, True
            self.current_time = self.total_time() + 1
            return
        if final_time <= 0:
            self.reset()
            return
        low, high = 0, self._ev_count
        while high - low > term_criterion:
            middle = (low + high) // 2
            probe = int(self._t[middle])
            if probe > final_time:
                high = middle
            elif probe < final_time:
                low = middle + 1
            else:
                self._pos = middle + 1
                self.current_time = final_time
                self.done = self._pos >= self._ev_count
                return middle
        index = bisect.bisect_left(self._t, final_time, low, high)
        self.seek_event(index)
        self.current_time = final_time
        self.done = self._pos >= self._ev_count
        return index

    def total_time(self):
        if not self._ev_count:
            return 0
        return int(self._t[-1])
