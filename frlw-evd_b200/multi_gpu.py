"""Multi-GPU driver: independent recordings sharded across the GPUs of one box.

Every recording carries its own state (TAF FIFO reset per file, ``generate_taf.py:155-158``;
SAE ``memory = None`` per file, ``generate_surfaceofactiveevents.py:145``), so the path
shards by recording with NO data-path collective.  One process per GPU
(``torchrun --nproc-per-node N -m frlw_evd_b200.multi_gpu ...``); the only collective is the
final reduction of a small statistics vector (NCCL on GPUs, gloo in the CPU tests).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m frlw_evd_b200.multi_gpu \\
        -rep taf -raw_dir R -label_dir L -target_dir T -dataset gen4
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist

STATS = ("recordings", "events", "windows", "bytes_written", "seconds")


def assign_recordings(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first: recordings (by event count, descending; ties by index)
    go to the least loaded rank.  Deterministic, so every rank computes the same plan."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(p) for p in plan]


def list_recordings(raw_dir: str, label_dir: str) -> List[Tuple[str, str, str, str, int]]:
    """``(mode, name, event file, label file, payload bytes)`` for every ``*_td.dat`` found,
    in a deterministic order (sorted directory listings)."""
    found = []
    for mode in ("train", "val", "test"):
        try:
            listing = sorted(os.listdir(os.path.join(raw_dir, mode)))
        except Exception:
            continue
        for fname in listing:
            if fname[-3:] != "dat":
                continue
            name = fname[:-7]
            path = os.path.join(raw_dir, mode, fname)
            found.append((mode, name, path, os.path.join(label_dir, mode, name + "_bbox.npy"), os.path.getsize(path)))
    return found


def reduce_stats(local: dict, device) -> dict:
    """Sum of the counters and max of the wall time over all ranks (the path's only
    collective)."""
    vec = torch.tensor([float(local.get(k, 0)) for k in STATS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        total = vec.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        slowest = vec[-1:].clone()
        dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
        total[-1] = slowest[0]
        vec = total
    return {k: float(v) for k, v in zip(STATS, vec.tolist())}


def encode_one(rep: str, dataset: str, mode: str, name: str, event_file: str, label_file: str, target_dir: str) -> dict:
    """Encode one recording with the chosen representation and write its files."""
    from . import generate_eventcountimage as eci
    from . import generate_eventvolume as evol
    from . import generate_surfaceofactiveevents as sae
    from . import generate_taf as taf
    from .io import npy_events_tools
    from .recordings import AsyncWriter, DeviceRecording, Geometry, PinnedRing

    geom = Geometry.for_dataset(dataset)
    labels = npy_events_tools.read_label_times(label_file)
    writer = AsyncWriter()              # files are written by a thread pool behind a ring of pinned buffers
    if rep == "taf":                    # host pipeline: H2D, kernels and D2H overlap
        rec = DeviceRecording(event_file, decode=False)
        windows = taf.encode_recording_to_files(rec, labels, name, mode, target_dir, geom, writer)
        writer.close()
        return {"events": rec.n_events, "windows": windows, "bytes_written": writer.bytes_written}
    rec = DeviceRecording(event_file)
    ring = PinnedRing(rec.events.device)
    windows = 0
    if rep == "event_volume":
        windows = evol.encode_recording_to_files(rec, labels, name, mode, target_dir, geom, ring, writer)
    elif rep in ("count_image", "sae"):
        if rep == "count_image":
            sizes = eci.windows_for(dataset)
            chunks, folders = eci.encode_chunks(rec, labels, geom, sizes), ["EventCountImage{0}".format(n) for n in sizes]
        else:
            chunks, folders = sae.encode_chunks(rec, labels, geom), ["SurfaceOfActiveEvents{0}".format(lam) for lam in sae.LAMDAS]
        for chunk_labels, u8 in chunks:
            def emit(host, names=[name + "_" + str(label) + ".npy" for label in chunk_labels]):
                return [writer.put(host[i, k], target_dir, folder, mode, fname)
                        for i, fname in enumerate(names) for k, folder in enumerate(folders)]
            ring.push(u8.contiguous(), emit)
            windows += len(chunk_labels)
    else:
        raise ValueError("unknown representation " + rep)
    ring.flush()
    writer.close()
    return {"events": rec.events.n, "windows": windows, "bytes_written": writer.bytes_written}


def run(recordings, encode: Callable[..., dict], device, rank: int, world: int) -> dict:
    """Encode this rank's share of ``recordings`` and reduce the statistics."""
    mine = assign_recordings([r[-1] for r in recordings], world)[rank]
    local = dict.fromkeys(STATS, 0)
    tick = time.time()
    for i in mine:
        stats = encode(recordings[i])
        local["recordings"] += 1
        for k in ("events", "windows", "bytes_written"):
            local[k] += stats.get(k, 0)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    local["seconds"] = time.time() - tick
    return reduce_stats(local, device)


def main(argv=None):
    ap = argparse.ArgumentParser(description="shard recordings over the GPUs of one box")
    ap.add_argument("-rep", default="taf", choices=["taf", "count_image", "sae", "event_volume"])
    ap.add_argument("-raw_dir", type=str)
    ap.add_argument("-label_dir", type=str)
    ap.add_argument("-target_dir", type=str)
    ap.add_argument("-dataset", type=str, default="gen4")
    args = ap.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("frlw_evd_b200.multi_gpu needs CUDA devices (no CPU fallback)")
    if world > 1:
        from .affinity import bind_to_device
        bind_to_device(local_rank)          # pinned buffers on the GPU's NUMA node
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries the JSON line only: NCCL honours NCCL_DEBUG_FILE above the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    recordings = list_recordings(args.raw_dir, args.label_dir)
    totals = run(recordings,
                 lambda r: encode_one(args.rep, args.dataset, r[0], r[1], r[2], r[3], args.target_dir),
                 device, rank, world)
    if rank == 0:
        totals["world_size"] = world
        totals["Mevents_per_s"] = totals["events"] / max(totals["seconds"], 1e-9) / 1e6
        print(json.dumps(totals))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
