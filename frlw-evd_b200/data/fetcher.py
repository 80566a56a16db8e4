"""Online (streaming) encoding -- drop-in for the reference's ``data/fetcher.py``.

Same classes, constructor arguments and ``fetch()`` return tuple.  The reference keeps the
batch's events in a host numpy array, selects the slice of every step with a boolean mask on
the host and uploads it (``data/fetcher.py:35-50``); here the events are uploaded ONCE, sorted
by time on the device, and every step is an index range found with two binary searches.
``to_volume`` is one of the ``frlw_evd_b200.data.sparse_ops`` encoders (same plugin signature,
``data/fetcher.py:53``).
"""
from __future__ import annotations

import time

import numpy as np
import torch


class fetcher:
    def __init__(self, events, shape, labels, timestamps, filenames, events_window, event_volume_bins, infer_time, to_volume,
                 device="cuda"):
        self.events_window_abin = infer_time
        self.events_window = events_window
        self.event_volume_bins = event_volume_bins
        self.shape = shape
        self.memory = None
        self.total_time = int(timestamps[0, 1] - timestamps[0, 0])
        self.iter = 0
        self.labels = labels
        self.timestamps = timestamps
        self.filenames = filenames
        self.finish = False
        self.to_volume = to_volume
        self.device = torch.device(device)
        ev = torch.as_tensor(events).to(self.device)
        # a stable sort by time keeps the reference's event order inside every step (a boolean mask
        # preserves the array order; steps are disjoint time ranges)
        order = torch.argsort(ev[..., 3], stable=True)
        self.events = ev[order].contiguous()
        self._t = self.events[..., 3].contiguous()

    def _range(self, lo, hi):
        """Events with ``lo <= t < hi`` (``lo = None``: from the start)."""
        t = self._t
        a = 0 if lo is None else int(torch.searchsorted(t, torch.tensor([lo], dtype=t.dtype, device=t.device))[0])
        b = int(torch.searchsorted(t, torch.tensor([hi], dtype=t.dtype, device=t.device))[0])
        return self.events[a:b]

    def getLabels(self, timestamps):
        max_labels = 80
        tol = self.events_window_abin / 2 - 1
        padded_labels = torch.zeros((len(self.timestamps), max_labels, self.labels.shape[1] - 1)).float().to(self.labels.device)
        for batch in range(len(self.timestamps)):
            timestamp = timestamps[batch]
            labels_ = self.labels[(self.labels[:, 0] == batch) & (self.labels[:, 6] + tol >= timestamp) &
                                  (self.labels[:, 6] - tol <= timestamp)]
            if len(labels_) == 0:
                return None
            assert max_labels >= len(labels_)
            padded_labels[batch, range(len(labels_))] = labels_[:, 1:].float()
        return padded_labels

    def fetch(self):
        if self.iter == 0:
            events = self._range(None, self.events_window)                               # t < events_window
            self.iter += self.events_window
        else:
            events = self._range(self.iter, self.iter + self.events_window_abin)         # iter <= t < iter + abin
            self.iter += self.events_window_abin
        if self.iter >= self.total_time:
            self.finish = True
        start = time.time()
        volume, self.memory = self.to_volume(events, len(self.timestamps), self.shape, self.iter, self.memory,
                                             self.events_window, self.event_volume_bins, self.events_window_abin)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        represent_time = time.time() - start
        if self.events_window == 60000000:
            timestamps = self.timestamps[..., 1]
        else:
            timestamps = self.timestamps[..., 0] + self.iter
        labels = self.getLabels(timestamps)
        return volume, labels, timestamps, self.filenames, represent_time


class fetcherTrain(fetcher):
    def getLabels(self, timestamps):
        labels = super().getLabels(timestamps)
        if labels is not None:
            return torch.cat([labels[:, :, 4:5], labels[:, :, :4]], dim=-1)
        return None


class fetcherVal(fetcher):
    pass
