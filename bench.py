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

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._halt.is_set():
                if not self._armed.wait(0.05):
                    continue
                region = self.regions[self._region]
                region["mhz"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                region["mask"] |= pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self._halt.wait(0.001)
        except Exception as exc:          # clocks are evidence, not a dependency
            self.error = repr(exc)

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
        print(json.dumps({
            "impl": "reference", "metric": "Mevents/s encoded (TAF K=8, 1MP)", "value": value, "unit": "Mevents/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "Mevents/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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

    # end to end through the public call with host buffers
    e2e = None
    if not args.no_e2e:
        from frlw_evd_b200 import generate_taf as gt
        from frlw_evd_b200.recordings import Geometry
        geom = Geometry((720, 1280), GRID, dev, coord_maps=maps)
        raw_host = torch.from_numpy(records.view(np.uint8)).pin_memory()
        pipe = gt.HostPipeline(geom, windows, K, ABIN, windows_per_chunk=12, device=dev)
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
        e2e = {"value": total_events / (e2e_ms * 1e-3) / 1e6, "unit": "Mevents/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(raw_host.numel()), "d2h_bytes_per_step": int(u8_host.numel()),
               "path": "pinned .dat bytes -> H2D -> decode -> taf_stream -> leaky uint8 [K,2,Ht,Wt] -> D2H, "
                       "12-window chunks on three streams (generate_taf.HostPipeline)"}

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
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
