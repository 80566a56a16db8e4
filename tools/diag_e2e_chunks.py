#!/usr/bin/env python
"""End-to-end step time of generate_taf.HostPipeline on the bench workload for several chunk schedules."""
import json
import os
import sys

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
n = len(windows)
schedules = {"uniform 12": 12, "ramp 1/2/4/8": gt.ramped_chunks(n, 8, 1), "ramp 1/2/4/8/16": gt.ramped_chunks(n, 16, 1),
             "ramp 1/2/4/8/16/24": [1, 2, 4, 8, 16] + [24] * 5 + [16, 8, 4, 2, 1] + [16],
             "ramp 1/2/4/8/16/32": gt.ramped_chunks(n, 32, 1), "ramp 1..48": [1, 2, 4, 8, 16, 32, 48, 32, 24, 16, 8, 4, 2, 1]}
out_host = None
for name, sched in schedules.items():
    pipe = gt.HostPipeline(geom, windows, bench.K, bench.ABIN, windows_per_chunk=sched, device=dev)
    if out_host is None:
        out_host = torch.empty(pipe.out_shape, dtype=torch.uint8).pin_memory()
    pipe.run(raw_host, out_host)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        pipe.run(raw_host, out_host)
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"schedule": name, "chunks": len(pipe.chunks), "ms_per_step": a.elapsed_time(b) / 3}), flush=True)
    del pipe
