#!/usr/bin/env python
"""Writes a seeded synthetic 'checkpoint' (<out>.npz of TF-named variables + an empty <out>.meta so that the
reference's demo.py existence check passes).  python tools/make_synthetic_ckpt.py --net res101 --classes 21 --anchors 9 --out /tmp/x.ckpt"""
import argparse
import _init_paths  # noqa: F401
import numpy as np
from tf_faster_rcnn_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="res101")
ap.add_argument("--classes", type=int, default=21)
ap.add_argument("--anchors", type=int, default=9)
ap.add_argument("--out", required=True)
a = ap.parse_args()
np.savez(a.out + ".npz", **synth.make(a.net, a.classes, a.anchors))
open(a.out + ".meta", "w").close()
print("wrote", a.out + ".npz")
