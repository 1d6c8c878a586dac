#!/usr/bin/env python
"""Re-score a finished test_net run from its detections.pkl (counterpart of the reference's tools/reval.py:24-75; same
arguments).  `--nms` first re-applies per-class NMS at cfg.TEST.NMS through model.test.apply_nms (GPU);
without it the script is CPU-only.

    python tools/reval.py <output_dir> --imdb voc_2007_test [--comp] [--nms] [--set DATA_DIR /data ...]"""
import argparse
import os
import pickle
import sys

import _init_paths  # noqa: F401

from model.config import cfg, cfg_from_list
from datasets.factory import get_imdb


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Re-evaluate results")
    p.add_argument("output_dir", nargs=1, type=str, help="results directory holding detections.pkl")
    p.add_argument("--imdb", dest="imdb_name", default="voc_2007_test", type=str)
    p.add_argument("--matlab", dest="matlab_eval", action="store_true")
    p.add_argument("--comp", dest="comp_mode", action="store_true")
    p.add_argument("--nms", dest="apply_nms", action="store_true")
    p.add_argument("--set", dest="set_cfgs", default=None, nargs=argparse.REMAINDER)
    if argv is None and len(sys.argv) == 1:
        p.print_help()
        sys.exit(1)
    return p.parse_args(argv)


def from_dets(imdb_name, output_dir, args):
    imdb = get_imdb(imdb_name)
    imdb.competition_mode(args.comp_mode)
    imdb.config["matlab_eval"] = args.matlab_eval
    with open(os.path.join(output_dir, "detections.pkl"), "rb") as f:
        dets = pickle.load(f)
    if args.apply_nms:
        from model.test import apply_nms
        print("Applying NMS to all detections")
        dets = apply_nms(dets, cfg.TEST.NMS)
    print("Evaluating detections")
    return imdb.evaluate_detections(dets, output_dir)


if __name__ == "__main__":
    a = parse_args()
    if a.set_cfgs:
        cfg_from_list(a.set_cfgs)
    from_dets(a.imdb_name, os.path.abspath(a.output_dir[0]), a)
