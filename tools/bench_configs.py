#!/usr/bin/env python
"""BASELINE.json configs 2-4 on one GPU (secondary to bench.py): device-resident events,
CUDA-event timing of the whole driver loop, one JSON line per measurement.

  python tools/bench_configs.py [--gen1-seconds 60]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from frlw_evd_b200 import generate_eventcountimage as g_eci  # noqa: E402
from frlw_evd_b200 import generate_surfaceofactiveevents as g_sae  # noqa: E402
from frlw_evd_b200 import generate_taf as g_taf  # noqa: E402
from frlw_evd_b200 import ops, synth  # noqa: E402
from frlw_evd_b200.io import PSEELoader  # noqa: E402
from frlw_evd_b200.recordings import Geometry  # noqa: E402


class Rec:            # a DeviceRecording built from arrays instead of a file
    def __init__(self, t, x, y, p, dev):
        self.loader = PSEELoader.from_records(synth.pack_dat_records(t, x, y, p))
        self.events = ops.EventStream.from_numpy(t, x, y, p, dev)


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        tick = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - tick)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gen1-seconds", type=float, default=60.0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak, _ = bench.load_peaks()

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    # ---- GEN1 240x304, 1 Mev/s (configs 2 and 3)
    dur = int(args.gen1_seconds * 1e6)
    t, x, y, p = synth.make_stream(240, 304, dur, 1e6, 1001)
    labels = synth.label_times(dur)
    rec = Rec(t, x, y, p, dev)
    geom = Geometry.for_dataset("gen1")
    n = len(t)
    HW = 240 * 304

    plan = g_taf.plan_windows(rec.loader, labels)
    wins = [w.as_tuple() for w in plan]
    bins = sum(w[3] for w in wins)
    for K in (8, 4):
        state = ops.taf_fresh_state(geom.grid, K, dev)
        out = torch.empty((len(wins), 2 * K, 240, 304), dtype=torch.float32, device=dev)
        s = wall(lambda: ops.taf_stream(rec.events, wins, 10000, geom.grid, K, state, None, False, out))
        algo = 9 * n + len(wins) * 8 * K * HW
        emit(config="3: TAF K=%d, GEN1 %gs @1 Mev/s, state carried" % (K, args.gen1_seconds), events=n, windows=len(wins),
             bins=bins, ms=s * 1e3, Mevents_per_s=n / s / 1e6, frac_of_measured_peak=algo / s / 1e9 / peak)

    s = wall(lambda: [None for _ in g_eci.encode_recording(rec, labels, geom, g_eci.windows_for("gen1"))], reps=1)
    emit(config="2: Event Count Image driver (N=50k/100k/200k per label), GEN1 %gs" % args.gen1_seconds, events=n,
         labels=len(labels), ms=s * 1e3, labels_per_s=len(labels) / s, Mevents_per_s=n / s / 1e6)
    s = wall(lambda: [None for _ in g_eci.encode_recording_stream(rec, labels, geom, g_eci.windows_for("gen1"))], reps=3)
    algo = 5 * n + len(labels) * 3 * (2 * HW + 2 * 256 * 320)       # events + uint8 count frame + uint8 image per window
    emit(config="2: Event Count Image driver, whole-stream kernel (128 labels per call), GEN1 %gs" % args.gen1_seconds,
         events=n, labels=len(labels), ms=s * 1e3, labels_per_s=len(labels) / s, Mevents_per_s=n / s / 1e6,
         frac_of_measured_peak=algo / s / 1e9 / peak)
    s = wall(lambda: [None for _ in g_sae.encode_recording(rec, labels, geom, "train")], reps=1)
    emit(config="2: SAE driver, one call per label (3 lambdas, memory carried), GEN1 %gs" % args.gen1_seconds, events=n,
         labels=len(labels), ms=s * 1e3, labels_per_s=len(labels) / s, Mevents_per_s=n / s / 1e6)
    s = wall(lambda: [None for _ in g_sae.encode_recording_stream(rec, labels, geom)], reps=3)
    algo = 9 * n + len(labels) * (8 * HW + 3 * 2 * 256 * 320)       # events + f32 frame + uint8 decays per label
    emit(config="2: SAE driver, whole-stream kernel (256 labels per call), GEN1 %gs" % args.gen1_seconds, events=n,
         labels=len(labels), ms=s * 1e3, labels_per_s=len(labels) / s, Mevents_per_s=n / s / 1e6,
         frac_of_measured_peak=algo / s / 1e9 / peak)
    # Event Volume driver (generate_eventvolume.py:118-169): 250 / 500 / 1000 ms windows per label, K = 5
    from frlw_evd_b200 import generate_eventvolume as g_ev
    plan = g_ev.label_windows(rec.loader, labels)
    splats = sum(hi - lo for _, ws in plan for lo, hi, _, _ in ws)
    s = wall(lambda: [None for _ in g_ev.encode_recording(rec, labels, geom)], reps=2)
    emit(config="V2: Event Volume driver (250/500/1000 ms per label, K=5, span kernels), GEN1 %gs" % args.gen1_seconds, events=n,
         labels=len(plan), window_events=splats, ms=s * 1e3, labels_per_s=len(plan) / s, window_Mevents_per_s=splats / s / 1e6)
    del rec

    # ---- 1MP (config 4): TAF K=4 next to the headline K=8, native-grid variant
    t, x, y, p = bench.get_stream(1002, 10.0, 1e7)
    n = len(t)
    records = synth.pack_dat_records(t, x, y, p)
    wins = bench.plan(records, 10.0)
    ev = ops.EventStream.from_numpy(t, x, y, p, dev)
    maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
    for K, grid, m in ((4, bench.GRID, maps), (8, bench.GRID, maps), (8, bench.SENSOR, None)):
        HWg = grid[0] * grid[1]
        nw = len(wins) if grid == bench.GRID else 40          # native grid: 59 MB per tensor, keep 40 windows
        state = ops.taf_fresh_state(grid, K, dev)
        out = torch.empty((nw, 2 * K, grid[0], grid[1]), dtype=torch.float32, device=dev)
        ww = wins[:nw]
        nev = sum(w[1] - w[0] for w in ww)
        s = wall(lambda: ops.taf_stream(ev, ww, 10000, grid, K, state, m, False, out))
        algo = 9 * nev + nw * 8 * K * HWg
        emit(config="4: TAF K=%d, 1MP on %dx%d grid" % (K, grid[0], grid[1]), events=nev, windows=nw, ms=s * 1e3,
             Mevents_per_s=nev / s / 1e6, frac_of_measured_peak=algo / s / 1e9 / peak)
        del out


    # Event Volume driver on the 1MP stream, gen4 policy (the windows of a label hold up to 10 M events)
    from frlw_evd_b200 import generate_eventvolume as g_ev
    rec = Rec(t, x, y, p, dev)
    geom4 = Geometry.for_dataset("gen4")
    labels4 = synth.label_times(10_000_000)
    plan = g_ev.label_windows(rec.loader, labels4)
    splats = sum(hi - lo for _, ws in plan for lo, hi, _, _ in ws)
    s = wall(lambda: [None for _ in g_ev.encode_recording(rec, labels4, geom4)], reps=2)
    emit(config="V2: Event Volume driver (250/500/1000 ms per label, K=5, span kernels), 1MP 10s gen4 policy", events=n,
         labels=len(plan), window_events=splats, ms=s * 1e3, labels_per_s=len(plan) / s, window_Mevents_per_s=splats / s / 1e6)


if __name__ == "__main__":
    main()
