#!/usr/bin/env python
"""GEN1 60 s recording: one warm pass and one profiled pass of the whole-stream Event Count Image and SAE drivers
(for `ncu --metrics gpu__time_duration.sum` launch lists and wall-clock splits)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frlw_evd_b200 import generate_eventcountimage as g_eci  # noqa: E402
from frlw_evd_b200 import generate_surfaceofactiveevents as g_sae  # noqa: E402
from frlw_evd_b200 import ops, synth  # noqa: E402
from frlw_evd_b200.io import PSEELoader  # noqa: E402
from frlw_evd_b200.recordings import Geometry  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda", 0)
dur = int(seconds * 1e6)
t, x, y, p = synth.make_stream(240, 304, dur, 1e6, 1001)
labels = synth.label_times(dur)


class Rec:
    def __init__(self):
        self.loader = PSEELoader.from_records(synth.pack_dat_records(t, x, y, p))
        self.events = ops.EventStream.from_numpy(t, x, y, p, dev)


rec, geom = Rec(), Geometry.for_dataset("gen1")
for name, fn in (("eci", lambda: [None for _ in g_eci.encode_chunks(rec, labels, geom, g_eci.windows_for("gen1"))]),
                 ("sae", lambda: [None for _ in g_sae.encode_chunks(rec, labels, geom)])):
    fn()
    torch.cuda.synchronize()
    tick = time.perf_counter()
    fn()
    host = time.perf_counter() - tick
    torch.cuda.synchronize()
    print("%s: host enqueue %.2f ms, until idle %.2f ms" % (name, host * 1e3, (time.perf_counter() - tick) * 1e3), flush=True)
