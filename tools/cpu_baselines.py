#!/usr/bin/env python
"""CPU baselines beside BASELINE.json configs 1-5: the reference's algorithm as restated by
the oracle (torch CPU ops, all host threads), on BOUNDED samples of the synthetic workloads -- driver runs include
the file decode, float64 staging and the file writes like the reference scripts.  One JSON line per config.

  python tools/cpu_baselines.py [--out profiles/r2_cpu_baselines.jsonl]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frlw_evd_b200 import synth  # noqa: E402
from oracle import drivers as od  # noqa: E402
from oracle import encoders as oe  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lines = []

    def emit(**kw):
        kw.update(cores=cores, torch_threads=torch.get_num_threads(), kind="port (oracle restatement of the reference)")
        lines.append(kw)
        print(json.dumps(kw), flush=True)

    # config 1: Event Volume K=8, one GEN1 50 ms window (~200 k events), best of 5
    t, x, y, p = synth.make_stream(240, 304, 50000, 4e6, 1000)
    ev = torch.from_numpy(np.stack([x, y, t / 50000.0, p], 1).astype(np.float64))
    best = min(_timed(lambda: oe.event_volume(ev.clone(), (240, 304), 8)) for _ in range(5))
    emit(config="1: Event Volume K=8, GEN1 one 50 ms window", events=len(t), seconds=best, Mevents_per_s=len(t) / best / 1e6,
         sample="the whole config, best of 5")

    def driver(tag, sensor, dataset, duration, rate, seed, runs):
        with tempfile.TemporaryDirectory() as tmp:
            (tt, _, _, _), labels = synth.write_recording(os.path.join(tmp, "raw"), os.path.join(tmp, "raw"), "test", "r", sensor,
                                                         duration, rate, seed)
            for name, fn in runs:
                s = _timed(lambda: fn(os.path.join(tmp, "raw"), os.path.join(tmp, "raw"), os.path.join(tmp, "out_" + name), dataset))
                emit(config=tag + name, events=len(tt), labels=len(labels), seconds=s, Mevents_per_s=len(tt) / s / 1e6,
                     sample="first %.1f s of the recording (%d events, %d labels), driver end to end incl. decode and file writes"
                            % (duration / 1e6, len(tt), len(labels)))

    torch.set_num_threads(1)       # the reference's SAE scatter races across CPU threads
    driver("2: GEN1 60 s @1 Mev/s, ", "gen1", "gen1", 6_000_000, 1e6, 1001, [("Surface of Active Events driver (1 thread)", od.run_sae)])
    torch.set_num_threads(cores)
    driver("2: GEN1 60 s @1 Mev/s, ", "gen1", "gen1", 6_000_000, 1e6, 1001, [("Event Count Image driver", od.run_count_image)])
    driver("3: GEN1 60 s @1 Mev/s, ", "gen1", "gen1", 6_000_000, 1e6, 1001, [("TAF K=8 driver, state carried", od.run_taf)])
    driver("4: 1MP @10 Mev/s, gen4 policy, ", "gen4", "gen4", 1_200_000, 1e7, 1002, [("TAF K=8 driver", od.run_taf)])
    driver("4: 1MP @10 Mev/s, gen4 policy, ", "gen4", "gen4", 450_000, 1e7, 1002, [("Event Volume driver (250/500/1000 ms, K=5)", od.run_event_volume)])
    driver("5: one of the 64 recordings (1MP, 2 s, 20 M events), ", "gen4", "gen4", 2_000_000, 1e7, 2000,
           [("decode + TAF K=8 driver; x64 recordings = extrapolated", od.run_taf)])
    if args.out:
        with open(args.out, "w") as fh:
            for line in lines:
                fh.write(json.dumps(line) + "\n")


def _timed(fn):
    tick = time.perf_counter()
    fn()
    return time.perf_counter() - tick


if __name__ == "__main__":
    main()
