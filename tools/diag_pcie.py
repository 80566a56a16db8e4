#!/usr/bin/env python
"""PCIe diagnostics for the end-to-end path: pinned H2D / D2H bandwidth alone and concurrently,
with the byte counts of the bench workload.  Prints one JSON line."""
import json
import torch

dev = torch.device("cuda", 0)
n_in, n_out = 800_000_000, 1_038_090_240
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best


def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


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


res = {"h2d_ms": timed(h2d), "d2h_ms": timed(d2h), "both_ms": timed(both), "both_9_chunks_ms": timed(chunked(9)),
       "both_40_chunks_ms": timed(chunked(40))}
res["h2d_GBs"] = n_in / res["h2d_ms"] / 1e6
res["d2h_GBs"] = n_out / res["d2h_ms"] / 1e6
print(json.dumps(res))
