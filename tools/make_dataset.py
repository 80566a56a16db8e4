#!/usr/bin/env python
"""Write a small synthetic dataset in the reference's directory layout
(<root>/{train,val,test}/<name>_td.dat + <name>_bbox.npy), e.g. for multi_gpu runs.

  python tools/make_dataset.py ROOT --sensor gen4 --recordings 8 --seconds 2 --rate 1e7
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frlw_evd_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("root")
ap.add_argument("--sensor", default="gen4")
ap.add_argument("--recordings", type=int, default=8)
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--rate", type=float, default=1e7)
args = ap.parse_args()
for r in range(args.recordings):
    synth.write_recording(args.root, args.root, "train", "rec%03d" % r, args.sensor, int(args.seconds * 1e6), args.rate, 1000 + r)
print("wrote", args.recordings, "recordings under", args.root)
