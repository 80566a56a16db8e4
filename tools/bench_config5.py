#!/usr/bin/env python
"""BASELINE config 5: a batch of synthetic 1MP recordings (T = 2 s, 10 Mev/s, 20 M events, seeds
1000 + r), `.dat` decode + TAF K=8, sharded by recording over the GPUs of one box with
`multi_gpu.assign_recordings`; the only collective is `multi_gpu.reduce_stats`.

Two numbers per run: decode + encode with the raw bytes already in device memory,
and end to end from pinned host bytes to pinned uint8 output (generate_taf.HostPipeline).
Times are CUDA-event times per rank, reduced with MAX; events are reduced with SUM.

    python tools/bench_config5.py --recordings 64                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/bench_config5.py --recordings 64
"""
import argparse
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def make(seed, seconds, rate):
    from frlw_evd_b200 import synth
    t, x, y, p = synth.make_stream(bench.SENSOR[0], bench.SENSOR[1], int(seconds * 1e6), rate, seed)
    return synth.pack_dat_records(t, x, y, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--recordings", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--rate", type=float, default=1e7)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to the CPUs next to its GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from frlw_evd_b200.affinity import bind_to_device
    placement = bind_to_device(local) if world > 1 and not args.no_bind else {"bound": False}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the JSON line only: NCCL honours NCCL_DEBUG_FILE above the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from frlw_evd_b200 import generate_taf as gt, multi_gpu, ops, synth
    from frlw_evd_b200.io import PSEELoader
    from frlw_evd_b200.recordings import Geometry

    sizes = [int(round(args.rate * args.seconds)) * 8] * args.recordings       # payload bytes, known up front
    mine = multi_gpu.assign_recordings(sizes, world)[rank]
    workers = max(1, min(8, (os.cpu_count() or 8) // world))
    with ProcessPoolExecutor(max_workers=workers) as pool:
        records = list(pool.map(make, [1000 + r for r in mine], [args.seconds] * len(mine), [args.rate] * len(mine)))

    maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
    geom = Geometry(bench.SENSOR, bench.GRID, dev, coord_maps=maps)
    labels = synth.label_times(int(args.seconds * 1e6))
    plans = [[w.as_tuple() for w in gt.plan_windows(PSEELoader.from_records(rec), labels)] for rec in records]
    n_events = sum(w[1] - w[0] for plan in plans for w in plan)
    n_windows = sum(len(plan) for plan in plans)

    # (A) raw bytes resident on the device: decode + bucketing + tile kernel per recording
    raw_dev = [torch.from_numpy(rec.view(np.uint8)).to(dev) for rec in records]
    soa = ops.EventStream.empty(max(len(r) for r in records), dev)
    out = torch.empty((max(len(p) for p in plans), 2 * bench.K, bench.GRID[0], bench.GRID[1]), dtype=torch.float32, device=dev)

    def resident():
        for raw, rec, plan in zip(raw_dev, records, plans):
            ev = soa.slice(0, len(rec))
            ops.decode_dat(raw, ev)
            state = ops.taf_fresh_state(bench.GRID, bench.K, dev)
            ops.taf_stream(ev, plan, bench.ABIN, bench.GRID, bench.K, state, maps, False, out[:len(plan)])

    # (B) end to end from pinned host memory
    raw_host = [torch.from_numpy(rec.view(np.uint8)).pin_memory() for rec in records]
    pipes = [gt.HostPipeline(geom, plan, bench.K, bench.ABIN, windows_per_chunk=12, device=dev) for plan in plans]
    u8_host = torch.empty(max(p.out_shape for p in pipes), dtype=torch.uint8).pin_memory()

    def end_to_end():
        for raw, pipe in zip(raw_host, pipes):
            pipe.state.fill_(-6000.0)                # every recording starts from a fresh FIFO
            pipe.run(raw, u8_host[:pipe.out_shape[0]])

    def timed(fn):
        fn()
        best = None
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        return best

    res_ms = timed(resident)
    e2e_ms = timed(end_to_end)
    stats_a = multi_gpu.reduce_stats({"recordings": len(mine), "events": n_events, "windows": n_windows,
                                      "bytes_written": 0, "seconds": res_ms * 1e-3}, dev)
    stats_b = multi_gpu.reduce_stats({"recordings": len(mine), "events": n_events, "windows": n_windows,
                                      "bytes_written": n_windows * u8_host[0].numel(), "seconds": e2e_ms * 1e-3}, dev)
    if rank == 0:
        print(json.dumps({
            "config": "5: %d 1MP recordings (%gs @ %g Mev/s), .dat decode + TAF K=8, sharded by recording" % (
                args.recordings, args.seconds, args.rate / 1e6),
            "n_gpus": world, "recordings": int(stats_a["recordings"]), "events": int(stats_a["events"]),
            "windows": int(stats_a["windows"]),
            "resident_ms": stats_a["seconds"] * 1e3, "resident_Mevents_per_s": stats_a["events"] / stats_a["seconds"] / 1e6,
            "e2e_ms": stats_b["seconds"] * 1e3, "e2e_Mevents_per_s": stats_b["events"] / stats_b["seconds"] / 1e6,
            "e2e_bytes_out": int(stats_b["bytes_written"]), "host_placement_rank0": placement,
            "timing": "CUDA events per rank, best of %d, max over ranks; events summed over ranks" % args.reps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
