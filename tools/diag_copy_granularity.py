#!/usr/bin/env python
"""One e2e step's bytes (1.04 GB device -> pinned host, 0.80 GB pinned host -> device) as 1 ... 100 copies per direction:
what the chunked pipeline pays for its copy granularity (about 50 us per extra copy).  python tools/diag_copy_granularity.py"""
import torch, time
dev = torch.device("cuda", 0)
nbytes = 198 * 8 * 2 * 512 * 640
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
hin = torch.empty(800_000_000, dtype=torch.uint8).pin_memory()
din = torch.empty(800_000_000, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(n_out, n_in, with_in=True):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        step = (nbytes + n_out - 1) // n_out
        for i in range(0, nbytes, step):
            h[i:i + step].copy_(d[i:i + step], non_blocking=True)
    if with_in:
        with torch.cuda.stream(s2):
            step = (800_000_000 + n_in - 1) // n_in
            for i in range(0, 800_000_000, step):
                din[i:i + step].copy_(hin[i:i + step], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
for n_out, n_in, w in [(1, 1, False), (25, 1, False), (1, 1, True), (25, 25, True), (8, 25, True), (4, 25, True), (50, 50, True), (100, 100, True)]:
    run(n_out, n_in, w)
    print("d2h in %3d copies, h2d in %3d copies (%s): %.2f ms" % (n_out, n_in, "both" if w else "d2h only", min(run(n_out, n_in, w) for _ in range(3))))
