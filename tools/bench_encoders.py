#!/usr/bin/env python
"""Secondary measurements (not the headline): per-window encoders and the decoder on the
1MP synthetic stream, CUDA-event timed, with the SURVEY 8d algorithmic bytes and the
fraction of the measured HBM peak.  One JSON line per encoder.

  python tools/bench_encoders.py [--seconds 2] [--iters 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from frlw_evd_b200 import ops, synth  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    peak, _ = bench.load_peaks()
    dev = torch.device("cuda", 0)
    t, x, y, p = bench.get_stream(1002, args.seconds, 1e7)
    n = len(t)
    ev = ops.EventStream.from_numpy(t, x, y, p, dev)
    maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
    H, W = bench.GRID
    HW = H * W
    edges = np.searchsorted(t, np.arange(0, int(args.seconds * 1e6) + 1, 50000))
    windows = [(int(edges[i]), int(edges[i + 1]), i * 50000) for i in range(len(edges) - 1)]
    nw = len(windows)

    def report(name, ms, algo_bytes, events, extra=None):
        line = {"encoder": name, "ms": ms, "Mevents_per_s": events / ms / 1e3, "algorithmic_bytes": algo_bytes,
                "achieved_GBs": algo_bytes / ms / 1e6, "frac_of_measured_peak": algo_bytes / ms / 1e6 / peak,
                "events": events, "windows": nw}
        line.update(extra or {})
        print(json.dumps(line))

    # decode: 8 B in, 9 B out per event
    raw = torch.from_numpy(synth.pack_dat_records(t, x, y, p).view(np.uint8)).to(dev)
    dec = ops.EventStream.empty(n, dev)
    report("decode_dat", timed(lambda: ops.decode_dat(raw, dec), args.iters), 17 * n, n)

    for K in (5, 8):
        out = torch.empty((2 * K, H, W), dtype=torch.float32, device=dev)

        def run_ev():
            for lo, hi, t0 in windows:
                ops.event_volume(ev.slice(lo, hi), t0, 50000, (H, W), K, maps, out)
        report("event_volume_K%d_50ms_windows" % K, timed(run_ev, args.iters), 9 * n + nw * 8 * K * HW, n)

    for K in (5, 8):
        outs = torch.empty((nw, 2 * K, H, W), dtype=torch.float32, device=dev)
        report("event_volume_stream_K%d_50ms_windows" % K,
               timed(lambda: ops.event_volume_stream(ev, windows, 50000, (H, W), K, maps, outs), args.iters),
               9 * n + nw * 8 * K * HW, n)

    # the same windows through the slice sort + span kernel (time-ordered streams; also serves nested windows)
    import bisect
    segments, spans = ops.plan_ev_spans([(lo, hi, t0, 50000) for lo, hi, t0 in windows], lambda i: int(t[i]),
                                        lambda T, lo, hi: bisect.bisect_left(t, T, lo, hi))
    for K in (5, 8):
        outs = torch.empty((nw, 2 * K, H, W), dtype=torch.float32, device=dev)
        pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ms = timed(lambda: ops.event_volume_spans(ev, segments, spans, (H, W), K, maps, outs, tile_events=pair), args.iters)
        torch.cuda.synchronize()
        report("event_volume_spans_K%d_50ms_windows" % K, ms, 9 * n + nw * 8 * K * HW, n,
               {"tile_kernel_ms": pair[0].elapsed_time(pair[1])})

    def run_eci():
        for lo, hi, _ in windows:
            ops.count_image(ev.slice(lo, hi), (H, W), maps)
    report("count_image_50ms_windows", timed(run_eci, args.iters), 5 * n + nw * 8 * HW, n)

    # whole-stream count image: the gen4 driver's nested last-N windows (400k / 800k / 1.2M events) per 50 ms label
    sizes = (400_000, 800_000, 1_200_000)
    eci_windows = [(max(hi - s, 0), hi) for _, hi, _ in windows for s in sizes]
    eci_out = torch.empty((len(eci_windows), 2, H, W), dtype=torch.uint8, device=dev)

    def run_eci_stream():
        ops.count_lut_u8_batch(ops.count_stream(ev, eci_windows, (H, W), maps), None, None, eci_out)
    report("count_stream_last_400k_800k_1200k_per_label_u8", timed(run_eci_stream, args.iters),
           5 * n + len(eci_windows) * 2 * 2 * HW, n, {"windows_encoded": len(eci_windows)})

    def run_eci_labels():
        for _, hi, _ in windows:
            ops.count_images_u8(ev.slice(max(hi - max(sizes), 0), hi), sizes, (H, W), (H, W), maps)
    report("count_per_label_last_400k_800k_1200k_u8", timed(run_eci_labels, args.iters),
           5 * n + len(eci_windows) * 2 * HW, n, {"windows_encoded": len(eci_windows)})

    # training-time read path (8f rank 2): a batch of 32 gen4 TAF samples, uint8 [16,512,640] each, augmented
    from frlw_evd_b200.data import dataset_gpu as dg
    import time as _time
    from oracle import dataset_read as dr
    rng = np.random.default_rng(5)
    vols_host = rng.integers(0, 256, (32, 16, H, W), dtype=np.uint8)
    vols = torch.from_numpy(vols_host).to(dev)
    aug = []
    for i in range(32):
        sr = 1.0 if i % 2 == 0 else float(rng.uniform(1.0, 1.5))
        cx = int(rng.uniform(int(W - sr * W), 0)) if sr > 1.0 else 0
        cy = int(rng.uniform(int(H - sr * H), 0)) if sr > 1.0 else 0
        aug.append((sr, cx, cy, bool(i & 2)))
    batch_out = torch.empty((32, 16, H, W), dtype=torch.float32, device=dev)
    ms = timed(lambda: dg.augment_batch(vols, (H, W), aug, batch_out), args.iters)
    tick = _time.perf_counter()
    for i in range(4):
        dr.augment_sample(vols_host[i].astype(np.float32), (H, W), *aug[i])
    cpu_ms = (_time.perf_counter() - tick) / 4 * 1e3
    print(json.dumps({"encoder": "dataset_read_batch32_gen4_taf", "ms": ms, "samples_per_s": 32 / ms * 1e3,
                      "algorithmic_bytes": 32 * 16 * HW * 5, "achieved_GBs": 32 * 16 * HW * 5 / ms / 1e6,
                      "frac_of_measured_peak": 32 * 16 * HW * 5 / ms / 1e6 / peak,
                      "cpu_oracle_ms_per_sample": cpu_ms, "cpu_threads": torch.get_num_threads()}))

    lam = [0.00001, 0.0000025, 0.000001]

    def run_sae():
        mem = None
        for lo, hi, t0 in windows:
            _, mem = ops.sae(ev.slice(lo, hi), (H, W), lam, mem, t0 + 50000, maps)
    report("sae_3lambda_50ms_windows", timed(run_sae, args.iters), 9 * n + nw * (8 * 3 * HW + 8 * HW), n)

    # whole-stream SAE: frames of latest timestamps (f32 [2,H,W] per window) + batched uint8 decays
    sae_windows = [(lo, hi, t0 + 50000, int(t[lo]) if hi > lo else 0, int(t[hi - 1]) if hi > lo else 0) for lo, hi, t0 in windows]
    frames = torch.empty((nw, 2, H, W), dtype=torch.float32, device=dev)
    u8 = torch.empty((nw, 3, 2, H, W), dtype=torch.uint8, device=dev)
    nows = [w[2] for w in sae_windows]

    def run_sae_stream():
        latest, _ = ops.sae_stream(ev, sae_windows, (H, W), None, maps, frames)
        ops.sae_decay_u8_batch(latest, nows, lam, None, None, u8)
    report("sae_stream_3lambda_50ms_windows_u8", timed(run_sae_stream, args.iters), 9 * n + nw * (8 * HW + 3 * 2 * HW), n)
    report("sae_stream_frames_only", timed(lambda: ops.sae_stream(ev, sae_windows, (H, W), None, maps, frames), args.iters),
           9 * n + nw * 8 * HW, n)


if __name__ == "__main__":
    main()
