#!/usr/bin/env python
"""Where the ordered TAF tile kernel spends its cycles, per role (diagnostic build only):
   EVREP_NVCC_EXTRA=-DEVREP_TAF_TIMING python -m frlw_evd_b200.build --force && python tools/diag_taf_timing.py
Prints, averaged over the tile CTAs, the cycles of the producer warp (total / waiting for a free stage) and of
worker thread 0 (total / waiting for a full stage / bin-end barriers / sweeps / accumulate / push)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from frlw_evd_b200 import _lib, ops, synth  # noqa: E402


def main():
    seconds, rate = 10.0, 1e7
    t, x, y, p = bench.get_stream(1002, seconds, rate)
    windows = bench.plan(synth.pack_dat_records(t, x, y, p), seconds)
    dev = torch.device("cuda", 0)
    ev = ops.EventStream.from_numpy(t, x, y, p, dev)
    maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
    state = ops.taf_fresh_state(bench.GRID, bench.K, dev)
    out = torch.empty((len(windows), 2 * bench.K, *bench.GRID), dtype=torch.float32, device=dev)
    for _ in range(3):
        ops.taf_stream(ev, windows, bench.ABIN, bench.GRID, bench.K, state, maps, False, out)
    torch.cuda.synchronize()
    lib = _lib.load()
    n_tiles = 400
    host = (ctypes.c_ulonglong * (8 * n_tiles))()
    rc = lib.evrep_debug_taf_timing(host, n_tiles)
    assert rc == 0, rc
    a = np.frombuffer(host, dtype=np.uint64).reshape(n_tiles, 8).astype(np.float64)
    a = a[a[:, 2] > 0]
    names = ["producer total", "producer wait free stage", "worker total", "worker wait full stage", "worker bin barriers",
             "worker sweeps", "worker accumulate", "worker push"]
    print("tiles", len(a))
    for i, n in enumerate(names):
        print("%-26s mean %10.0f  min %10.0f  max %10.0f cycles" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))


if __name__ == "__main__":
    main()
