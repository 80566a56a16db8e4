"""Temporal Active Focus -- drop-in for the reference's ``generate_taf.py``.

Same names and signatures: ``taf_cuda``, ``generate_taf_cuda``, ``leaky_transform`` and the
``-raw_dir -label_dir -target_dir -dataset`` command line (``python -m
frlw_evd_b200.generate_taf ...``).  The functions run the one-bin CUDA kernel on the
reference's float64 staging layout; the command line plans all windows of a recording on
the host (``plan_windows``, ``generate_taf.py:160-187`` of the reference) and runs them
through the bucketing pass + persistent tile kernel (``ops.taf_stream``).
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass
from typing import List

import torch

from . import ops
from .recordings import DeviceRecording, Geometry, dump_u8, iter_recordings, parse_args

ABIN = 10000          # generate_taf.py:99  (delta tau = 10 ms)
VOLUME_BINS = 8       # :100
MIN_EVENT_COUNT = 50000000   # :92


# ------------------------------------------------------------------ reference functions
def taf_cuda(x, y, t, p, shape, volume_bins, past_volume):
    """``generate_taf.py:19-58``: ``x, y, p`` integer tensors, ``t`` float32 in [0, 1].
    Returns ``(f32 [2K,H,W], f32 state [H,W,2,K], seconds)``."""
    events = torch.stack([x.double(), y.double(), t.double(), p.double()], dim=1)
    return generate_taf_cuda(events, shape, past_volume, volume_bins)


def generate_taf_cuda(events, shape, past_volume=None, volume_bins=5):
    """``generate_taf.py:60-67``: ``events`` float64 ``[n, >=4]`` (x, y, t_norm, p[, z])."""
    tick = time.time()
    out, state = ops.taf_bin_aos64(events, tuple(shape), int(volume_bins), past_volume)
    torch.cuda.synchronize()
    return out, state, time.time() - tick


def leaky_transform(ecd):
    """``generate_taf.py:69-76``: ``255 * max(0, 1 - log1p(-v) / 8.7)``."""
    return ops.leaky_transform(ecd)


# ------------------------------------------------------------------------ window plans
@dataclass
class TafWindow:
    label: int
    fresh: bool
    start_time: int
    end_time: int
    start_count: int
    end_count: int

    def n_bins(self, abin=ABIN) -> int:
        return math.ceil((self.end_time - self.start_time) / abin)

    def as_tuple(self, abin=ABIN):
        return (self.start_count, self.end_count, self.start_time, self.n_bins(abin), int(self.fresh))


def plan_windows(loader, labels, abin=ABIN, volume_bins=VOLUME_BINS, min_event_count=MIN_EVENT_COUNT) -> List[TafWindow]:
    """All windows of one recording (``generate_taf.py:155-187,237-238``).  A window is
    *fresh* (state reset, own aligned start) when it starts after the previous window's
    end; otherwise it continues from the previous end and its end is snapped to the
    10 ms grid (Python ``round``: half to even) and clamped to the last timestamp."""
    span = abin * volume_bins
    t_upper, c_upper = -1e16, -1
    plan = []
    for label in labels:
        end_time = int(label)
        end_count = loader.seek_time(end_time)
        if end_count is None:
            continue
        loader.seek_event(max(end_count - min_event_count, 0))
        start_time = int(loader.current_time)
        if end_time - start_time < span:
            start_time = end_time - span
        else:
            start_time = end_time - round((end_time - start_time - span) / abin) * abin - span
        fresh = start_time > t_upper
        if fresh:
            start_count = loader.seek_time(start_time)
            if start_count is None or start_time < 0:
                start_count = 0
        else:
            start_count, start_time = c_upper, t_upper
            end_time = round((end_time - start_time) / abin) * abin + start_time
            end_time = min(end_time, loader.total_time())
            end_count = loader.seek_time(end_time)
        plan.append(TafWindow(label, fresh, int(start_time), int(end_time), int(start_count), int(end_count)))
        t_upper, c_upper = end_time, end_count
    return plan


def encode_recording(rec: DeviceRecording, plan: List[TafWindow], geom: Geometry, abin=ABIN,
                     volume_bins=VOLUME_BINS, windows_per_launch=32):
    """Yield ``(label, u8 [K,2,Ht,Wt])`` for every planned window: slot 0 = newest bin,
    first half -> ``bins{K/2}``, second half -> ``bins{K}`` (``generate_taf.py:226-235``)."""
    grid = geom.grid
    state = ops.taf_fresh_state(grid, volume_bins, rec.events.device)
    for lo in range(0, len(plan), windows_per_launch):
        chunk = plan[lo:lo + windows_per_launch]
        volumes = ops.taf_stream(rec.events, [w.as_tuple(abin) for w in chunk], abin, grid, volume_bins,
                                 state, geom.coord_maps)
        for w, vol in zip(chunk, volumes):
            yield w.label, ops.taf_leaky_u8(vol, volume_bins, geom.target, geom.resize_maps)


def ramped_chunks(n_windows: int, peak: int = 12, first: int = 3):
    """Chunk sizes for ``HostPipeline``: small chunks at both ends (the device->host copy starts sooner and the
    last chunk's tail is short), ``peak`` windows in between."""
    head, size = [], first
    while size < peak and sum(head) + size <= n_windows // 2:
        head.append(size)
        size *= 2
    body = n_windows - 2 * sum(head)
    return head + [peak] * (body // peak) + ([body % peak] if body % peak else []) + head[::-1]


class HostPipeline:
    """End-to-end TAF encoding of one recording held in HOST memory: pinned ``.dat`` payload
    in, pinned uint8 ``[n_windows, K, 2, Ht, Wt]`` out (the bytes of the ``bins*`` files).

    The windows are processed in chunks; three streams overlap the host->device copy of
    chunk c+1, the kernels of chunk c (decode, bucketing, tile kernel, fused
    leaky/flip/resize/uint8 epilogue) and the device->host copy of chunk c-1.  The FIFO state
    is carried between chunks in a device tensor."""

    def __init__(self, geom: Geometry, plan, K=VOLUME_BINS, abin=ABIN, windows_per_chunk=None, device="cuda"):
        self.geom, self.K, self.abin, self.device = geom, K, abin, torch.device(device)
        self.windows = [w if isinstance(w, tuple) else w.as_tuple(abin) for w in plan]
        # `windows_per_chunk`: one size for all chunks, or the list of chunk sizes; default: 1, 2, 4 windows at both ends
        # and 8 in between (measured best on the bench workload: 23.1 ms against 23.6 for uniform 12-window chunks)
        if windows_per_chunk is None:
            windows_per_chunk = ramped_chunks(len(self.windows), 8, 1)
        sizes = ([windows_per_chunk] * -(-len(self.windows) // windows_per_chunk) if isinstance(windows_per_chunk, int)
                 else list(windows_per_chunk))
        self.chunks, i = [], 0
        for size in sizes:
            if i >= len(self.windows):
                break
            self.chunks.append((i, min(i + size, len(self.windows))))
            i += size
        if i < len(self.windows):
            self.chunks.append((i, len(self.windows)))
        max_ev = max((self.windows[b - 1][1] - self.windows[a][0] for a, b in self.chunks), default=0)
        max_w = max((b - a for a, b in self.chunks), default=0)
        H, W = geom.grid
        Ht, Wt = geom.target
        dev = self.device
        self.raw = [torch.empty(max(max_ev, 1) * 8, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.soa = [ops.EventStream.empty(max(max_ev, 1), dev) for _ in range(2)]
        # a .dat payload is ordered in time (the reference's loader seeks by bisection on that assumption), so the one-pass
        # kernels may be selected: EVREP_TAF_PATH=ordered (bin-major sort) or =sliced (shared-memory tile kernel, which
        # writes the file bytes itself when no resize follows).  The default is the general two-pass bucketing.
        self.assume_ordered = os.environ.get("EVREP_TAF_PATH", "") in ("ordered", "sliced")
        self.fused_u8 = os.environ.get("EVREP_TAF_PATH", "") == "sliced" and tuple(geom.grid) == tuple(geom.target)
        self.vol = None if self.fused_u8 else torch.empty((max(max_w, 1), 2 * K, H, W), dtype=torch.float32, device=dev)
        self.violations = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ring = None                   # pinned chunk buffers of the `sink` mode, allocated on first use
        self.u8 = [torch.empty((max(max_w, 1), K, 2, Ht, Wt), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.state = ops.taf_fresh_state(geom.grid, K, dev)
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.out_shape = (len(self.windows), K, 2, Ht, Wt)

    def run(self, raw_host: torch.Tensor, out_host: torch.Tensor = None, sink=None) -> torch.Tensor:
        """``raw_host``: pinned uint8 payload of the whole recording (8 bytes/event).  The uint8 tensors go either
        to ``out_host`` (pinned uint8 tensor of shape ``self.out_shape``) or, chunk by chunk, to
        ``sink(first_window, last_window, host_array) -> futures``: then only a ring of three pinned chunk
        buffers is page-locked, whatever the number of windows, and a buffer is reused once the futures its
        ``sink`` call returned (the file writes) are done."""
        assert (out_host is None) != (sink is None), "give either out_host or sink"
        ev = torch.cuda.Event
        in_done, raw_free, comp_done, out_done = ([ev() for _ in range(2)] for _ in range(4))
        start = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_stream(start)
        if sink is not None and self.ring is None:
            per = self.u8[0][0].numel()
            self.ring = [dict(buf=torch.empty(self.u8[0].shape[0] * per, dtype=torch.uint8, pin_memory=True), done=None,
                              span=None, futures=[]) for _ in range(3)]

        def retire(slot):
            if slot["span"] is not None:
                slot["done"].synchronize()
                a, b = slot["span"]
                n = (b - a) * self.u8[0][0].numel()
                slot["futures"] = list(sink(a, b, slot["buf"][:n].numpy().reshape((b - a,) + tuple(self.out_shape[1:]))) or [])
                slot["span"] = None

        for c, (a, b) in enumerate(self.chunks):
            k = c & 1
            e0, e1 = self.windows[a][0], self.windows[b - 1][1]
            n = e1 - e0
            with torch.cuda.stream(self.s_in):
                if c >= 2:
                    self.s_in.wait_event(raw_free[k])
                self.raw[k][:n * 8].copy_(raw_host[e0 * 8:e1 * 8], non_blocking=True)
                in_done[k].record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(in_done[k])
                soa = self.soa[k].slice(0, n)
                # a .dat payload is ordered in time (the reference's loader seeks by bisection on that
                # assumption); the ordered kernels count violations, `order_violations` reports them
                soa.ordered = self.assume_ordered
                ops.decode_dat(self.raw[k][:n * 8], soa)
                raw_free[k].record(self.s_comp)
                local = [(w[0] - e0, w[1] - e0, w[2], w[3], w[4]) for w in self.windows[a:b]]
                if c >= 2:
                    self.s_comp.wait_event(out_done[k])
                if self.fused_u8:
                    # no resize between grid and target: the tile kernel writes the file bytes itself
                    ops.taf_stream(soa, local, self.abin, self.geom.grid, self.K, self.state, self.geom.coord_maps,
                                   out_u8=self.u8[k][:b - a], want_f32=False)
                else:
                    vol = self.vol[:b - a]
                    ops.taf_stream(soa, local, self.abin, self.geom.grid, self.K, self.state, self.geom.coord_maps, False, vol)
                    ops.taf_leaky_u8_batch(vol, self.K, self.geom.target, self.geom.resize_maps, self.u8[k][:b - a])
                if self.assume_ordered:
                    status = ops.order_violations_tensor(self.device)
                    self.violations += status[0] + status[2]
                comp_done[k].record(self.s_comp)
            if sink is not None:
                if c:
                    retire(self.ring[(c - 1) % 3])           # the previous chunk has had a whole chunk of kernels to land
                slot = self.ring[c % 3]
                retire(slot)
                for f in slot["futures"]:
                    f.result()                               # its files are written: the buffer is free
                slot["futures"] = []
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(comp_done[k])
                if sink is None:
                    out_host[a:b].copy_(self.u8[k][:b - a], non_blocking=True)
                else:
                    slot["buf"][:(b - a) * self.u8[0][0].numel()].copy_(self.u8[k][:b - a].reshape(-1), non_blocking=True)
                out_done[k].record(self.s_out)
                if sink is not None:
                    slot["done"], slot["span"] = ev(), (a, b)
                    slot["done"].record(self.s_out)
        for s in (self.s_in, self.s_comp, self.s_out):
            start.wait_stream(s)
        if sink is not None:
            for slot in self.ring:
                retire(slot)
            for slot in self.ring:
                for f in slot["futures"]:
                    f.result()
                slot["futures"] = []
        return out_host

    def order_violations(self) -> int:
        """Events found outside the bin their position implies, over all runs so far (synchronises).
        Non-zero means the payload was not ordered in time (or the sort could not run to completion) and the run
        has to be repeated with ``EVREP_TAF_PATH=bucketed``."""
        return int(self.violations.item())


def encode_recording_to_files(rec: DeviceRecording, labels, name: str, mode: str, target_dir: str, geom: Geometry,
                              writer, volume_bins=VOLUME_BINS, abin=ABIN) -> int:
    """Host pipeline + file layout of ``generate_taf.py:226-235``: pinned ``.dat`` payload in,
    ``taf/<mode>/bins{K/2}`` and ``bins{K}`` files out, chunk by chunk through a ring of three pinned
    buffers (the page-locked memory does not grow with the number of labels).  Returns the number of windows."""
    plan = plan_windows(rec.loader, labels, abin, volume_bins)
    if not plan:
        return 0
    pipe = HostPipeline(geom, plan, volume_bins, abin)
    half = volume_bins // 2

    def sink(a, b, host):
        futures = []
        for i, w in enumerate(plan[a:b]):
            fname = name + "_" + str(w.label) + ".npy"
            futures.append(writer.put(host[i, :half], target_dir, "taf", mode, "bins{0}".format(half), fname))
            futures.append(writer.put(host[i, half:], target_dir, "taf", mode, "bins{0}".format(volume_bins), fname))
        return futures
    pipe.run(rec.raw_pinned, sink=sink)
    return len(plan)


def main(argv=None):
    from .recordings import AsyncWriter
    args = parse_args("gen4", argv)
    geom = Geometry.for_dataset(args.dataset)
    writer = AsyncWriter()
    total_time, total_count = 0.0, 0
    for mode, name, event_file, labels in iter_recordings(args.raw_dir, args.label_dir):
        rec = DeviceRecording(event_file, decode=False)
        tick = time.time()
        n = encode_recording_to_files(rec, labels, name, mode, args.target_dir, geom, writer)
        if mode == "test":
            total_time += time.time() - tick
            total_count += n
    writer.close()
    if total_count:
        print("Average Representation time: ", total_time / total_count)


if __name__ == "__main__":
    main()
