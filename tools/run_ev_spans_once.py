#!/usr/bin/env python
"""One warm call + `--iters` calls of the span Event Volume (K=8, 1MP stream, 50 ms windows) for ncu captures."""
import argparse
import bisect
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from frlw_evd_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--K", type=int, default=8)
args = ap.parse_args()
dev = torch.device("cuda", 0)
t, x, y, p = bench.get_stream(1002, args.seconds, 1e7)
ev = ops.EventStream.from_numpy(t, x, y, p, dev)
maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
edges = list(range(0, int(args.seconds * 1e6) + 1, 50000))
windows = [(bisect.bisect_left(t, a), bisect.bisect_left(t, b), a, 50000) for a, b in zip(edges[:-1], edges[1:])]
segments, spans = ops.plan_ev_spans(windows, lambda i: int(t[i]), lambda T, lo, hi: bisect.bisect_left(t, T, lo, hi))
out = torch.empty((len(windows), 2 * args.K, *bench.GRID), dtype=torch.float32, device=dev)
for _ in range(args.iters + 1):
    ops.event_volume_spans(ev, segments, spans, bench.GRID, args.K, maps, out)
torch.cuda.synchronize()
print("ok", len(windows), "windows", len(segments), "segments")
