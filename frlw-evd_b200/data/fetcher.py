"""Online (streaming) encoding -- the role of the reference's ``data/fetcher.py`` (same three class
names, constructor arguments, ``fetch()`` tuple, ``finish`` / ``iter`` / ``memory`` attributes), built
for a device-resident stream:

* the batch's events are uploaded ONCE and ordered by time on the device; the index range of every
  step -- the reference selects it with a boolean mask over the whole host array and uploads the
  selection, step after step (``data/fetcher.py:35-50``) -- comes from a table of cut points computed
  for all steps by one ``searchsorted`` when the fetcher is built;
* ``fetch()`` does not wait for the GPU: the encoder call is bracketed by CUDA events and
  ``represent_time`` is read from them later (``represent_times()``); pass ``sync_timing=True`` for the
  reference's blocking per-step number;
* the labels are bucketed per batch sample once, so ``getLabels`` is two bisections per sample.

``to_volume`` is any encoder with the plugin signature of ``frlw_evd_b200.data.sparse_ops``
(``data/fetcher.py:53``), e.g. ``generate_taf_online_cuda``, whose FIFO state lives in ``memory`` on
the device from step to step.
"""
from __future__ import annotations

import bisect
import time

import numpy as np
import torch

MAX_LABELS = 80          # data/fetcher.py:23
WHOLE_RECORDING = 60000000   # data/fetcher.py:57: with a 60 s window the label time is the end of the recording


class _TimeOrderedEvents:
    """Rows ``(b, x, y, t, p)`` of a batch on the device, ordered by ``t`` (stable: rows with equal
    timestamps keep the order a boolean mask over the original array would give them)."""

    def __init__(self, events, device):
        rows = torch.as_tensor(events).to(device)
        order = torch.argsort(rows[..., 3], stable=True)
        self.rows = rows[order].contiguous()
        self.times = self.rows[..., 3].contiguous()

    def cuts(self, edges):
        """Index of the first row with ``t >= edge`` for every edge (one device call, one read back)."""
        probe = torch.as_tensor(np.asarray(edges), dtype=self.times.dtype, device=self.times.device)
        return torch.searchsorted(self.times, probe).cpu().tolist()


class _LabelBuckets:
    """Label rows ``(b, ..., t at column 6)`` grouped by batch sample and ordered by time."""

    def __init__(self, labels, n_samples):
        self.labels = labels
        host = labels.detach().cpu().numpy()
        self.rows, self.times = [], []
        for b in range(n_samples):
            idx = np.nonzero(host[:, 0] == b)[0]
            idx = idx[np.argsort(host[idx, 6], kind="stable")]
            self.rows.append(idx)
            self.times.append(host[idx, 6])

    def around(self, sample, when, tol):
        """Rows of ``sample`` with ``|t - when| <= tol``, in the order of the label array."""
        times = self.times[sample]
        a, b = bisect.bisect_left(times, when - tol), bisect.bisect_right(times, when + tol)
        return np.sort(self.rows[sample][a:b])


class fetcher:
    def __init__(self, events, shape, labels, timestamps, filenames, events_window, event_volume_bins, infer_time, to_volume,
                 device="cuda", sync_timing=False):
        self.shape, self.labels, self.timestamps, self.filenames = shape, labels, timestamps, filenames
        self.events_window, self.events_window_abin, self.event_volume_bins = events_window, infer_time, event_volume_bins
        self.to_volume = to_volume
        self.memory, self.iter, self.finish = None, 0, False
        self.total_time = int(timestamps[0, 1] - timestamps[0, 0])
        self.device = torch.device(device)
        self.sync_timing = sync_timing
        self._stream = _TimeOrderedEvents(events, self.device)
        self.events = self._stream.rows
        self._buckets = None
        # step k covers [clock[k], clock[k + 1]): the first one everything before `events_window`
        self._clock = [None, events_window]
        while self._clock[-1] < self.total_time:
            self._clock.append(self._clock[-1] + infer_time)
        self._cut = [0] + self._stream.cuts(self._clock[1:])
        self._step = 0
        self._events_pairs = []

    def _advance(self):
        if self._step + 1 >= len(self._clock):           # fetch() past the end of the recording: extend the table
            self._clock.append(self._clock[-1] + self.events_window_abin)
            self._cut.append(self._stream.cuts([self._clock[-1]])[0])
        lo, hi = self._cut[self._step], self._cut[self._step + 1]
        self._step += 1
        self.iter = self._clock[self._step]
        self.finish = self.iter >= self.total_time
        return self._stream.rows[lo:hi]

    def getLabels(self, timestamps):
        """``data/fetcher.py:22-33``: per sample the labels within half a step of its timestamp, zero padded to
        ``[B, 80, columns - 1]``; ``None`` as soon as one sample has none."""
        if self._buckets is None:
            self._buckets = _LabelBuckets(self.labels, len(self.timestamps))
        tol = self.events_window_abin / 2 - 1
        padded = torch.zeros((len(self.timestamps), MAX_LABELS, self.labels.shape[1] - 1), dtype=torch.float32,
                             device=self.labels.device)
        for sample in range(len(self.timestamps)):
            rows = self._buckets.around(sample, float(timestamps[sample]), tol)
            if len(rows) == 0:
                return None
            assert len(rows) <= MAX_LABELS
            padded[sample, :len(rows)] = self.labels[torch.as_tensor(rows, device=self.labels.device)][:, 1:].float()
        return padded

    def fetch(self):
        events = self._advance()
        on_gpu = self.device.type == "cuda"
        tick = time.time()
        if on_gpu and not self.sync_timing:
            begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            begin.record()
        volume, self.memory = self.to_volume(events, len(self.timestamps), self.shape, self.iter, self.memory,
                                             self.events_window, self.event_volume_bins, self.events_window_abin)
        if on_gpu and not self.sync_timing:
            end.record()
            self._events_pairs.append((begin, end))
        elif on_gpu:
            torch.cuda.synchronize()
        represent_time = time.time() - tick
        if self.events_window == WHOLE_RECORDING:
            timestamps = self.timestamps[..., 1]
        else:
            timestamps = self.timestamps[..., 0] + self.iter
        return volume, self.getLabels(timestamps), timestamps, self.filenames, represent_time

    def represent_times(self):
        """Device time of every ``to_volume`` call so far in seconds (synchronises once)."""
        if self._events_pairs:
            self._events_pairs[-1][1].synchronize()
        return [a.elapsed_time(b) * 1e-3 for a, b in self._events_pairs]


class fetcherTrain(fetcher):
    def getLabels(self, timestamps):
        """Training wants (class, box) instead of (box, class) (``data/fetcher.py:64-70``)."""
        labels = super().getLabels(timestamps)
        return None if labels is None else torch.cat([labels[..., 4:5], labels[..., :4]], dim=-1)


class fetcherVal(fetcher):
    pass
