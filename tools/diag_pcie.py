#!/usr/bin/env python
"""Host <-> device copy calibration for the end-to-end path, with the byte counts of one bench step (0.8 GB in,
1.04 GB out per GPU): pinned H2D alone, D2H alone, both at once, both in chunks; with a write-combined pinned
source for the H2D leg as a variant.  Under torchrun every rank copies at the same time (barrier before each
measurement) and rank 0 prints ONE JSON line with the slowest rank's times and the aggregate bandwidths:

  python tools/diag_pcie.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/diag_pcie.py
"""
import ctypes
import json
import os

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)

n_in, n_out = 800_000_000, 1_038_090_240
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def write_combined(n):
    """Pinned, write-combined host buffer (cudaHostAllocWriteCombined): not snooped by the CPU caches."""
    rt = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n), ctypes.c_uint(4)) != 0:
        return None
    buf = (ctypes.c_uint8 * n).from_address(ptr.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            m = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            ms = float(m[0])
        best = ms if best is None else min(best, ms)
    return best


def h2d(src=None):
    src = h_in if src is None else src
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_in.copy_(src, non_blocking=True)


def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def chunked(parts):
    def run():
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        for i in range(parts):
            a, b = n_in * i // parts, n_in * (i + 1) // parts
            c, d = n_out * i // parts, n_out * (i + 1) // parts
            with torch.cuda.stream(s1):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[c:d].copy_(d_out[c:d], non_blocking=True)
    return run


res = {"n_gpus": world, "h2d_bytes": n_in, "d2h_bytes": n_out,
       "h2d_ms": timed(h2d), "d2h_ms": timed(d2h), "both_ms": timed(both), "both_17_chunks_ms": timed(chunked(17))}
wc = write_combined(n_in)
if wc is not None:
    res["h2d_write_combined_ms"] = timed(lambda: h2d(wc))
    res["both_write_combined_ms"] = timed(lambda: (h2d(wc), d2h()))
res["h2d_GBs_aggregate"] = world * n_in / res["h2d_ms"] / 1e6
res["d2h_GBs_aggregate"] = world * n_out / res["d2h_ms"] / 1e6
res["both_GBs_aggregate"] = world * (n_in + n_out) / res["both_ms"] / 1e6
res["cpus"] = os.cpu_count()
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
