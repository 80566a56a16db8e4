#!/usr/bin/env python
"""Timeline of generate_taf.HostPipeline on the bench workload: per chunk, when the host->device
copy, the kernels and the device->host copy start and end (ms from the start of the step)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from frlw_evd_b200 import generate_taf as gt, ops, synth  # noqa: E402
from frlw_evd_b200.recordings import Geometry  # noqa: E402

dev = torch.device("cuda", 0)
t, x, y, p = bench.get_stream(1002, 10.0, 1e7)
records = synth.pack_dat_records(t, x, y, p)
windows = bench.plan(records, 10.0)
maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
geom = Geometry((720, 1280), bench.GRID, dev, coord_maps=maps)
raw_host = torch.from_numpy(records.view(np.uint8)).pin_memory()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pipe = gt.HostPipeline(geom, windows, bench.K, bench.ABIN, windows_per_chunk=per, device=dev)
out_host = torch.empty(pipe.out_shape, dtype=torch.uint8).pin_memory()
for _ in range(2):
    pipe.run(raw_host, out_host)
torch.cuda.synchronize()

marks = []
orig_copy = torch.Tensor.copy_


def mark(stream, name, c):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    marks.append((name, c, e, time.perf_counter()))


# re-implementation of HostPipeline.run with event marks
def run():
    ev = torch.cuda.Event
    in_done, raw_free, comp_done, out_done = ([ev() for _ in range(2)] for _ in range(4))
    start = torch.cuda.current_stream(dev)
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record(start)
    host0 = time.perf_counter()
    for s in (pipe.s_in, pipe.s_comp, pipe.s_out):
        s.wait_stream(start)
    for c, (a, b) in enumerate(pipe.chunks):
        k = c & 1
        e0, e1 = pipe.windows[a][0], pipe.windows[b - 1][1]
        n = e1 - e0
        with torch.cuda.stream(pipe.s_in):
            if c >= 2:
                pipe.s_in.wait_event(raw_free[k])
            mark(pipe.s_in, "h2d_begin", c)
            pipe.raw[k][:n * 8].copy_(raw_host[e0 * 8:e1 * 8], non_blocking=True)
            in_done[k].record(pipe.s_in)
            mark(pipe.s_in, "h2d_end", c)
        with torch.cuda.stream(pipe.s_comp):
            pipe.s_comp.wait_event(in_done[k])
            mark(pipe.s_comp, "comp_begin", c)
            soa = pipe.soa[k].slice(0, n)
            ops.decode_dat(pipe.raw[k][:n * 8], soa)
            raw_free[k].record(pipe.s_comp)
            local = [(w[0] - e0, w[1] - e0, w[2], w[3], w[4]) for w in pipe.windows[a:b]]
            vol = pipe.vol[:b - a]
            ops.taf_stream(soa, local, pipe.abin, pipe.geom.grid, pipe.K, pipe.state, pipe.geom.coord_maps, False, vol)
            mark(pipe.s_comp, "taf_end", c)
            if c >= 2:
                pipe.s_comp.wait_event(out_done[k])
            ops.taf_leaky_u8_batch(vol, pipe.K, pipe.geom.target, pipe.geom.resize_maps, pipe.u8[k][:b - a])
            comp_done[k].record(pipe.s_comp)
            mark(pipe.s_comp, "comp_end", c)
        with torch.cuda.stream(pipe.s_out):
            pipe.s_out.wait_event(comp_done[k])
            mark(pipe.s_out, "d2h_begin", c)
            out_host[a:b].copy_(pipe.u8[k][:b - a], non_blocking=True)
            out_done[k].record(pipe.s_out)
            mark(pipe.s_out, "d2h_end", c)
    for s in (pipe.s_in, pipe.s_comp, pipe.s_out):
        start.wait_stream(s)
    torch.cuda.synchronize()
    return t0, host0


t0, host0 = run()
end = torch.cuda.Event(enable_timing=True)
end.record()
torch.cuda.synchronize()
print("windows_per_chunk", per, "chunks", len(pipe.chunks), "total_ms %.2f" % t0.elapsed_time(end))
rows = {}
for name, c, e, host in marks:
    rows.setdefault(c, {})[name] = (t0.elapsed_time(e), (host - host0) * 1e3)
print("chunk  h2d[b,e]        comp[b, taf_e, e]          d2h[b,e]        host_enqueue_ms")
for c in sorted(rows):
    r = rows[c]
    print("%3d  %6.2f %6.2f   %6.2f %6.2f %6.2f   %6.2f %6.2f   %6.2f" % (
        c, r["h2d_begin"][0], r["h2d_end"][0], r["comp_begin"][0], r["taf_end"][0], r["comp_end"][0],
        r["d2h_begin"][0], r["d2h_end"][0], r["d2h_end"][1]))
