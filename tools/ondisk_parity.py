#!/usr/bin/env python
"""On-disk parity report: the four command-line drivers of this package against the oracle drivers
(restatements of the reference scripts, pinned by `tests/golden/drivers_digest.json`) on two small
synthetic datasets (gen1 and gen4 policy).  For every representation: files, bytes, bytes that differ
and the largest difference.  Writes `profiles/ondisk_parity.json`.

  python tools/ondisk_parity.py [--out profiles/ondisk_parity.json]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frlw_evd_b200 import generate_eventcountimage, generate_eventvolume, generate_surfaceofactiveevents, generate_taf, synth  # noqa: E402
from oracle import drivers as od  # noqa: E402

MAINS = {"count_image": generate_eventcountimage.main, "sae": generate_surfaceofactiveevents.main,
         "event_volume": generate_eventvolume.main, "taf": generate_taf.main}
ORACLE = {"count_image": od.run_count_image, "sae": od.run_sae, "event_volume": od.run_event_volume, "taf": od.run_taf}


def files_under(root):
    out = {}
    for folder, _, names in os.walk(root):
        for n in names:
            out[os.path.relpath(os.path.join(folder, n), root)] = os.path.join(folder, n)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ondisk_parity.json"))
    args = ap.parse_args()
    report = {"what": "bytes written by the generate_* command lines vs the oracle drivers on synthetic datasets",
              "datasets": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for dataset, duration, rate, seed in (("gen1", 1_500_000, 4e5, 5), ("gen4", 700_000, 3e6, 6)):
            raw, lab = os.path.join(tmp, dataset, "raw"), os.path.join(tmp, dataset, "labels")
            for mode, rec_seed in (("train", seed), ("test", seed + 100)):
                synth.write_recording(raw, lab, mode, "rec%d" % rec_seed, dataset, duration, rate, rec_seed)
            rows = {}
            for rep in MAINS:
                mine, ref = os.path.join(tmp, dataset, "mine_" + rep), os.path.join(tmp, dataset, "ref_" + rep)
                MAINS[rep](["-raw_dir", raw, "-label_dir", lab, "-target_dir", mine, "-dataset", dataset])
                ORACLE[rep](raw, lab, ref, dataset)
                a, b = files_under(mine), files_under(ref)
                n_bytes = diff = worst = 0
                for name in sorted(b):
                    x, y = np.fromfile(a[name], dtype=np.uint8), np.fromfile(b[name], dtype=np.uint8)
                    assert x.shape == y.shape, name
                    d = np.abs(x.astype(np.int16) - y.astype(np.int16))
                    n_bytes += x.size
                    diff += int((d != 0).sum())
                    worst = max(worst, int(d.max()) if d.size else 0)
                rows[rep] = {"files": len(b), "same_file_set": sorted(a) == sorted(b), "bytes": n_bytes, "bytes_that_differ": diff,
                             "flip_rate": diff / max(n_bytes, 1), "largest_difference": worst}
                print(dataset, rep, rows[rep], flush=True)
            report["datasets"][dataset] = rows
    with open(args.out, "w") as fh:
        json.dump(report, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main()
