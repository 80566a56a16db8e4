#!/usr/bin/env python
"""Where a register-resident TAF tile kernel spends its cycles, per role (diagnostic build only):
   EVREP_NVCC_EXTRA=-DEVREP_TILE_TIMING python -m frlw_evd_b200.build --force && python tools/diag_tile_timing.py
Averages over the tile CTAs the cycle counters of accumulate thread 0 and consumer thread 0.  The counters are compiled
into the packed-accumulator variant (taf_tile_pk_kernel, selected here through EVREP_TAF_TILE_KERNEL=pk); the figures of the
default kernel in profiles/r2_tile_timing.txt came from the same macros placed temporarily in taf_tile_ws_kernel."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from frlw_evd_b200 import _lib, ops, synth  # noqa: E402


def main():
    os.environ["EVREP_TAF_TILE_KERNEL"] = "pk"
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
    n_tiles = 147
    host = (ctypes.c_ulonglong * (16 * n_tiles))()
    rc = lib.evrep_debug_tile_timing(host, n_tiles)
    assert rc == 0, rc
    a = np.frombuffer(host, dtype=np.uint64).reshape(n_tiles, 16).astype(np.float64)
    names = {0: "accumulate total", 1: "acc: wait record chunks", 2: "acc: preload + barrier", 3: "acc: wait accumulator buffer",
             4: "acc: atomics", 5: "acc: batch feed", 8: "consumer total", 9: "cons: wait FULL", 10: "cons: read + clear",
             11: "cons: update", 12: "cons: wait staging tile", 13: "cons: staging"}
    print("tiles", len(a), "bins", sum(w[3] for w in windows))
    for i, n in names.items():
        print("%-30s mean %10.0f  min %10.0f  max %10.0f cycles" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))


if __name__ == "__main__":
    main()
