#!/usr/bin/env python
"""Tile-kernel variants of the TAF stream path at several event rates: `EVREP_TAF_TILE_KERNEL` = (default) / ws / pk.
   python tools/taf_tile_variants.py [rate ...]        (events per second over 1 MP, default 1e7 3e7 1e8; 2 s of stream)
Prints the step and the tile-kernel time (CUDA events, median of 10) for every (rate, variant)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from frlw_evd_b200 import ops, synth  # noqa: E402


def main():
    rates = [float(a) for a in sys.argv[1:]] or [1e7, 3e7, 1e8]
    dev = torch.device("cuda", 0)
    maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
    for rate in rates:
        seconds = 10.0 if rate <= 1e7 else 2.0
        t, x, y, p = bench.get_stream(1002, seconds, rate)
        windows = bench.plan(synth.pack_dat_records(t, x, y, p), seconds)
        ev = ops.EventStream.from_numpy(t, x, y, p, dev)
        out = torch.empty((len(windows), 2 * bench.K, *bench.GRID), dtype=torch.float32, device=dev)
        ref = None
        for variant in ("ws", "pk"):
            os.environ["EVREP_TAF_TILE_KERNEL"] = variant
            step, tile = [], []
            for it in range(13):
                state = ops.taf_fresh_state(bench.GRID, bench.K, dev)
                pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.taf_stream(ev, windows, bench.ABIN, bench.GRID, bench.K, state, maps, False, out, tile_events=pair)
                b.record()
                torch.cuda.synchronize()
                if it >= 3:
                    step.append(a.elapsed_time(b)); tile.append(pair[0].elapsed_time(pair[1]))
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(ref, out))
            print("rate %.0e events %d windows %d variant %s step %.3f ms tile %.3f ms equal_to_ws %s" % (
                rate, len(t), len(windows), variant, statistics.median(step), statistics.median(tile), same), flush=True)
    os.environ.pop("EVREP_TAF_TILE_KERNEL", None)


if __name__ == "__main__":
    main()
