#!/usr/bin/env python
"""Headline benchmark: Mevents/s encoded, TAF K=8 on a synthetic 1MP stream.

Workload (BASELINE.json configs[3]): one synthetic 1280x720
recording per GPU, 10 s at 10 Mev/s = 100 M events (seed 1002 + rank), label every 50 ms
from 100 ms (198 windows, 10 + 197 x 5 bins of 10 ms), Temporal Active Focus K=8 on the
reference's 512x640 grid (gen4 coordinate policy), float32 [2K,H,W] tensor + state written
per window.  A step = one pass over the whole recording: bucketing (6 kernels) + the
persistent tile kernel.  Inputs (900 MB of SoA events) exceed the 126 MB L2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`   : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`     : same metric through the public call with HOST buffers: pinned .dat bytes ->
            H2D -> decode -> TAF -> uint8 [K,2,Ht,Wt] tensors (the on-disk payload) -> D2H.
`roofline`: the tile kernel, timed live with CUDA events recorded around it by the library.
`cpu_baseline` / `--impl reference`: the reference's CPU algorithm (oracle port, torch CPU
            ops, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SENSOR = (720, 1280)
GRID = (512, 640)
K = 8
ABIN = 10000
CACHE = os.environ.get("EVREP_BENCH_CACHE", "/tmp/evrep_bench")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def get_stream(seed, seconds, rate):
    """Synthetic recording (SURVEY.md §8d generator), cached on local disk."""
    from frlw_evd_b200 import synth
    os.makedirs(CACHE, exist_ok=True)
    tag = os.path.join(CACHE, "s%d_%g_%g" % (seed, seconds, rate))
    names = [tag + "_%s.npy" % k for k in "txyp"]
    if all(os.path.isfile(n) for n in names):
        return tuple(np.load(n) for n in names)
    arrays = synth.make_stream(SENSOR[0], SENSOR[1], int(seconds * 1e6), rate, seed)
    for n, a in zip(names, arrays):
        np.save(n + ".tmp.npy", a)
        os.replace(n + ".tmp.npy", n)
    return arrays


def plan(records, seconds):
    from frlw_evd_b200 import generate_taf as gt
    from frlw_evd_b200 import synth
    from frlw_evd_b200.io import PSEELoader
    labels = synth.label_times(int(seconds * 1e6))
    return [w.as_tuple() for w in gt.plan_windows(PSEELoader.from_records(records), labels)]


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled with NVML while a timed region is armed.
    The thread (and NVML) is brought up before the warm-up so that the first sample
    lands inside the timed region, not after it."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.max_mhz, self._halt, self._armed = index, None, threading.Event(), threading.Event()
        self.regions, self._region, self.error = {}, None, None
        self.ready = threading.Event()          # NVML is up: samples start within a millisecond of arm()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.ready.set()
            while not self._halt.is_set():
                if not self._armed.wait(0.05):
                    continue
                region = self.regions[self._region]
                region["mhz"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                region["mask"] |= pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self._halt.wait(0.001)
        except Exception as exc:          # clocks are evidence, not a dependency
            self.error = repr(exc)
            self.ready.set()

    def arm(self, region):
        self.regions.setdefault(region, {"mhz": [], "mask": 0})
        self._region = region
        self._armed.set()

    def disarm(self):
        self._armed.clear()

    def stop(self):
        self._halt.set()
        self._armed.set()
        self.join(timeout=2)

    def report(self, region):
        r = self.regions.get(region, {"mhz": [], "mask": 0})
        return {"sm_mhz": statistics.median(r["mhz"]) if r["mhz"] else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if r["mask"] & bit], "samples": len(r["mhz"])}


def cpu_reference_leg(t, x, y, p, windows, sample_windows, repeats=1):
    """The reference's CPU algorithm (oracle port: same ATen ops, all host threads) on the
    first `sample_windows` windows.  Returns (Mevents/s, events, seconds, threads)."""
    import torch
    from oracle import drivers as od
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = windows[:sample_windows]
    hi = sample[-1][1]
    staged = torch.from_numpy(np.stack([x[:hi], y[:hi], t[:hi], p[:hi]], 1).astype(np.float64))
    n_ev = sum(w[1] - w[0] for w in sample)
    best = None
    for _ in range(repeats):
        tick = time.perf_counter()
        od.taf_windows_in_memory(staged, sample, ABIN, GRID, K, scale=(GRID[1] / SENSOR[1], GRID[0] / SENSOR[0]))
        dt = time.perf_counter() - tick
        best = dt if best is None else min(best, dt)
    return n_ev / best / 1e6, n_ev, best, threads


def cpu_reference_driver_leg(t, x, y, p, windows, sample_windows):
    """the reference DRIVER end to end on the host -- .dat file on disk, loader decode,
    float64 staging, the encode loops, leaky transform, flip, uint8, file writes (`generate_taf.py:78-243`, oracle
    port `oracle.drivers.run_taf`) -- for the first `sample_windows` labels.  Returns (Mevents/s, events, seconds)."""
    import tempfile
    import torch
    from frlw_evd_b200 import synth
    from oracle import drivers as od
    torch.set_num_threads(os.cpu_count() or 1)
    sample = windows[:sample_windows]
    hi = sample[-1][1]
    n_ev = sum(w[1] - w[0] for w in sample)
    labels = synth.label_times(int(t[hi - 1]) + 2)[:sample_windows]
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "raw", "test"))
        synth.write_dat(os.path.join(tmp, "raw", "test", "rec_td.dat"), t[:hi + 1], x[:hi + 1], y[:hi + 1], p[:hi + 1], *SENSOR)
        synth.write_bbox_npy(os.path.join(tmp, "raw", "test", "rec_bbox.npy"), labels)
        tick = time.perf_counter()
        od.run_taf(os.path.join(tmp, "raw"), os.path.join(tmp, "raw"), os.path.join(tmp, "out"), "gen4")
        dt = time.perf_counter() - tick
    return n_ev / dt / 1e6, n_ev, dt


def parity_check(ev, windows, t, x, y, p, maps, n_check=3):
    """Outside every timed region: the first windows of the benchmark's own workload (a fresh 10-bin window and two
    incremental 5-bin ones) through the kernels that are timed, against the CPU oracle.  The oracle is the checker only."""
    import torch
    from frlw_evd_b200 import ops
    from oracle import drivers as od
    sample = windows[:n_check]
    hi = sample[-1][1]
    staged = torch.from_numpy(np.stack([x[:hi], y[:hi], t[:hi], p[:hi]], 1).astype(np.float64))
    want, want_state = od.taf_windows_in_memory(staged, sample, ABIN, GRID, K, scale=(GRID[1] / SENSOR[1], GRID[0] / SENSOR[0]))
    state = ops.taf_fresh_state(GRID, K, ev.device)
    got = ops.taf_stream(ev, sample, ABIN, GRID, K, state, maps).cpu()
    worst = 0.0
    ok = True
    for i in range(n_check):
        a, b = got[i].numpy(), want[i].numpy()
        ok = ok and bool(np.allclose(a, b, rtol=1e-5, atol=1e-6))
        worst = max(worst, float(np.max(np.abs(a - b) / (1e-6 + 1e-5 * np.abs(b)))))
    ok = ok and bool(np.allclose(state.cpu().numpy(), want_state.numpy(), rtol=1e-5, atol=1e-6))
    return {"ok": ok, "windows": n_check, "events": int(sum(w[1] - w[0] for w in sample)), "tolerance": "1e-5 rel + 1e-6 abs",
            "worst_error_over_tolerance": worst, "against": "oracle.drivers.taf_windows_in_memory (CPU, float32 like the reference)"}


def ev_record(ev, t, seconds, maps, out, peak, steps, barrier):
    """The second north-star encoder on the same stream: Event Volume K=8, 50 ms windows back to back, through the span
    kernels (slice sort + tile kernel); device-resident, CUDA-event timed."""
    import bisect
    import torch
    from frlw_evd_b200 import ops
    edges = list(range(0, int(seconds * 1e6) + 1, 50000))
    wins = [(bisect.bisect_left(t, a), bisect.bisect_left(t, b), a, 50000) for a, b in zip(edges[:-1], edges[1:])]
    segments, spans = ops.plan_ev_spans(wins, lambda i: int(t[i]), lambda T, lo, hi: bisect.bisect_left(t, T, lo, hi))
    n_ev = sum(w[1] - w[0] for w in wins)
    need = len(wins) * 2 * K * GRID[0] * GRID[1]
    vol = (out.view(-1)[:need] if out.numel() >= need else torch.empty(need, dtype=torch.float32, device=ev.device)).view(len(wins), 2 * K, *GRID)
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for _ in range(3):
        ops.event_volume_spans(ev, segments, spans, GRID, K, maps, vol)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    for i in range(steps):
        ops.event_volume_spans(ev, segments, spans, GRID, K, maps, vol, tile_events=pairs[i])
    b.record()
    barrier()
    ms = a.elapsed_time(b) / steps
    tile_ms = statistics.mean(x.elapsed_time(y) for x, y in pairs)
    algo = 9 * n_ev + len(wins) * 4 * 2 * K * GRID[0] * GRID[1]
    return {"metric": "Mevents/s encoded (Event Volume K=8, 1MP)", "value": n_ev / (ms * 1e-3) / 1e6, "unit": "Mevents/s",
            "ms_per_step": ms, "windows": len(wins), "events": n_ev, "kernel": "ev_slice_tile_kernel", "kernel_ms": tile_ms,
            "algorithmic_bytes_per_launch": algo, "frac": algo / (tile_ms * 1e-3) / 1e9 / peak,
            "step_frac": algo / (ms * 1e-3) / 1e9 / peak,
            "path": "evrep_event_volume_spans: bins by bisection + slice sort + tile kernel (float32 [2K,H,W] per window)"}


def copy_floor(raw_host, u8_host, dev, barrier, reps=3):
    """The same bytes as one e2e step, nothing but the copies: pinned -> device and device -> pinned at the same time
    on two streams, every rank at once.  The e2e step cannot be faster than this."""
    import torch
    d_in = torch.empty(raw_host.numel(), dtype=torch.uint8, device=dev)
    d_out = torch.empty(u8_host.numel(), dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    flat_out = u8_host.view(-1)

    def once():
        cur = torch.cuda.current_stream(dev)
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(raw_host, non_blocking=True)
        with torch.cuda.stream(s2):
            flat_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    once()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    for _ in range(reps):
        once()
    b.record()
    barrier()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=10.0, help="recording length (default: the 100 M-event workload)")
    ap.add_argument("--rate", type=float, default=1e7)
    ap.add_argument("--cpu-windows", type=int, default=120, help="windows in the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "TAF K=8, 1MP 1280x720 -> 512x640 grid, %gs @ %g Mev/s per GPU, 50 ms windows" % (args.seconds, args.rate / 1e6)
    config = {"workload": workload, "events_per_gpu": int(round(args.rate * args.seconds)), "grid": list(GRID), "K": K,
              "abin_us": ABIN, "l2_policy": "inputs (SoA events, 9 B/event) larger than L2; outputs 4 GB per step"}

    if args.impl == "reference":
        if rank != 0:
            return
        t, x, y, p = get_stream(1002, args.seconds, args.rate)
        from frlw_evd_b200 import synth
        windows = plan(synth.pack_dat_records(t, x, y, p), args.seconds)
        for _ in range(max(args.warmup, 0) and 1):
            cpu_reference_leg(t, x, y, p, windows, 2)
        vals = []
        tick = time.perf_counter()
        for _ in range(args.steps):
            v, n_ev, dt, threads = cpu_reference_leg(t, x, y, p, windows, args.cpu_windows)
            vals.append(v)
        total = time.perf_counter() - tick
        value = statistics.mean(vals)
        sample = "first %d windows (%d events) of the workload per step, encoder loops of generate_taf.py:195-222" % (args.cpu_windows, n_ev)
        # the driver end to end (file decode, float64 staging, encode, leaky, uint8, file writes)
        drv_windows = max(2, min(args.cpu_windows, 24))
        drv_value, drv_events, drv_seconds = cpu_reference_driver_leg(t, x, y, p, windows, drv_windows)
        print(json.dumps({
            "impl": "reference", "metric": "Mevents/s encoded (TAF K=8, 1MP)", "value": value, "unit": "Mevents/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "Mevents/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": drv_value, "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "what": "the reference driver end to end on the host (BASELINE.md section 3 ii): .dat file -> loader decode -> "
                            "float64 staging -> encode loops -> leaky / flip / uint8 -> files; first %d labels (%d events), "
                            "%.1f s" % (drv_windows, drv_events, drv_seconds)},
        }))
        return

    import torch
    import torch.distributed as dist
    from frlw_evd_b200 import ops, synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    from frlw_evd_b200.affinity import bind_to_device
    placement = bind_to_device(local) if world > 1 else {"bound": False, "how": "single process"}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the JSON line only: NCCL honours NCCL_DEBUG_FILE above the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    t, x, y, p = get_stream(1002 + rank, args.seconds, args.rate)
    n_events = len(t)
    records = synth.pack_dat_records(t, x, y, p)
    windows = plan(records, args.seconds)
    nw = len(windows)
    n_in_windows = sum(w[1] - w[0] for w in windows)
    HW = GRID[0] * GRID[1]

    ev = ops.EventStream.from_numpy(t, x, y, p, dev)
    maps = ops.make_coord_maps(SENSOR, GRID, dev)
    state = ops.taf_fresh_state(GRID, K, dev)
    out = torch.empty((nw, 2 * K, GRID[0], GRID[1]), dtype=torch.float32, device=dev)

    def step(events=None):
        ops.taf_stream(ev, windows, ABIN, GRID, K, state, maps, False, out, events)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    check = parity_check(ev, windows, t, x, y, p, maps) if rank == 0 else None

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.ready.wait(10.0)
    barrier()
    sampler.arm("device")
    start.record()
    for i in range(args.steps):
        step(pairs[i])
    stop.record()
    barrier()
    sampler.disarm()
    ms = start.elapsed_time(stop) / args.steps
    tile_ms = statistics.mean(a.elapsed_time(b) for a, b in pairs)

    stats = torch.tensor([ms, float(n_in_windows), tile_ms], dtype=torch.float64, device=dev)
    if world > 1:          # the path's only collective: the final statistics reduction
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, total_events, tile_ms = float(mx[0]), float(sm[1]), float(mx[2])
    else:
        total_events = float(n_in_windows)
    value = total_events / (ms * 1e-3) / 1e6

    # roofline of the dominant kernel (the tile kernel).  Algorithmic bytes: events read
    # (9 B each) + one f32 [2K,H,W] tensor written per window + the state written once.
    # SURVEY.md §8d also counts a state write per window; the kernel keeps the state in
    # registers between windows, so those bytes are not moved and are NOT counted here.
    peak, peak_src = load_peaks()
    algo_bytes = 9 * n_in_windows + nw * (4 * 2 * K * HW) + 4 * 2 * K * HW
    survey_bytes = 9 * n_in_windows + nw * 2 * (4 * 2 * K * HW)
    achieved = algo_bytes / (tile_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "taf_tile_ws_kernel<8,6>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": tile_ms,
                "step_frac": algo_bytes / (ms * 1e-3) / 1e9 / peak,
                "note": "state kept on chip between windows: per-window state writes of SURVEY 8d are not moved and not counted",
                "frac_if_survey_formula_were_used": survey_bytes / (tile_ms * 1e-3) / 1e9 / peak}
    traffic_file = os.path.join(ROOT, "profiles", "taf_tile_traffic.json")
    if os.path.isfile(traffic_file):
        with open(traffic_file) as fh:
            roofline["traffic"] = json.load(fh).get("dram_bytes_per_launch")

    ev_rec = ev_record(ev, t, args.seconds, maps, out, peak, max(2, min(args.steps, 10)), barrier)

    # end to end through the public call with host buffers
    e2e = None
    if not args.no_e2e:
        from frlw_evd_b200 import generate_taf as gt
        from frlw_evd_b200.recordings import Geometry
        geom = Geometry((720, 1280), GRID, dev, coord_maps=maps)
        raw_host = torch.from_numpy(records.view(np.uint8)).pin_memory()
        pipe = gt.HostPipeline(geom, windows, K, ABIN, device=dev)
        u8_host = torch.empty(pipe.out_shape, dtype=torch.uint8).pin_memory()

        def e2e_step():
            pipe.run(raw_host, u8_host)

        e2e_step()
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(2, min(args.steps, 3))
        sampler.arm("e2e")
        s2.record()
        for _ in range(n_e2e):
            e2e_step()
        e2.record()
        barrier()
        sampler.disarm()
        e2e_ms = s2.elapsed_time(e2) / n_e2e
        if world > 1:
            tmax = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_ms = float(tmax[0])
        floor_ms = copy_floor(raw_host, u8_host, dev, barrier)
        if world > 1:
            tmax = torch.tensor([floor_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            floor_ms = float(tmax[0])
        e2e = {"value": total_events / (e2e_ms * 1e-3) / 1e6, "unit": "Mevents/s", "ms_per_step": e2e_ms,
               "copy_floor_ms": floor_ms, "copy_floor_note": "the step's bytes, both directions at once, all ranks at once, no kernels",
               "h2d_bytes_per_step": int(raw_host.numel()), "d2h_bytes_per_step": int(u8_host.numel()),
               "path": "pinned .dat bytes -> H2D -> decode -> taf_stream -> leaky uint8 [K,2,Ht,Wt] -> D2H, "
                       "1/2/4/8-window chunks on three streams (generate_taf.HostPipeline)"}

    sampler.stop()
    clocks = sampler.report("device")
    clocks["e2e_region"] = sampler.report("e2e")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, n_ev, dt, threads = cpu_reference_leg(t, x, y, p, windows, args.cpu_windows)
        cpu = {"value": v, "unit": "Mevents/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": "first %d windows (%d events) of the same stream, encoder loops of generate_taf.py:195-222 "
                         "(oracle port, torch CPU ops)" % (args.cpu_windows, n_ev)}

    # kernels of one device-resident step: chunk origins, count, two scans, tile bits, scatter,
    # tile kernel, plus the window / batch tables uploaded as kernel arguments (3840 B a launch)
    a16 = lambda v: (v + 15) // 16 * 16
    n_batches = sum(max(1, -(-w[3] // 16)) for w in windows)
    meta_bytes = 3 * a16(8 * nw) + a16(4 * nw) + a16(4 * (nw + 1)) + a16(16 * n_batches)
    launches_per_step = 7 + -(-meta_bytes // 3840)

    if rank == 0:
        print(json.dumps({
            "metric": "Mevents/s encoded (TAF K=8, 1MP)", "value": value, "unit": "Mevents/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "windows": nw, "events_in_windows_per_gpu": n_in_windows, "host_placement": placement,
            "parity_check": check, "ev": ev_rec,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
