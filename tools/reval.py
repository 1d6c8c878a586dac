#!/usr/bin/env python
"""Score an existing test_net run again from the `detections.pkl` it left in its output directory -- the command line of
the reference's tools/reval.py:24-75 (positional output dir, --imdb, --matlab, --comp, --nms), plus `--set K V ...` for cfg
overrides such as DATA_DIR.

    python tools/reval.py output/default/voc_2007_test/default --imdb voc_2007_test --comp

With --nms every class is first put through model.test.apply_nms at cfg.TEST.NMS (that step needs the GPU); otherwise
the script only touches the CPU."""
import argparse
import os
import pickle
import sys

import _init_paths  # noqa: F401

from datasets.factory import get_imdb
from model.config import cfg, cfg_from_list

FLAGS = (("--matlab", "matlab_eval", "evaluate with the MATLAB devkit (not provided here)"),
         ("--comp", "comp_mode", "competition mode: fixed result-file names, files are kept"),
         ("--nms", "apply_nms", "re-apply per-class NMS before scoring"))


def build_parser():
    parser = argparse.ArgumentParser(description="Re-evaluate results")
    parser.add_argument("output_dir", nargs=1, type=str, help="directory that holds detections.pkl")
    parser.add_argument("--imdb", dest="imdb_name", type=str, default="voc_2007_test", help="dataset to score against")
    for flag, dest, text in FLAGS:
        parser.add_argument(flag, dest=dest, action="store_true", help=text)
    parser.add_argument("--set", dest="set_cfgs", nargs=argparse.REMAINDER, default=None, help="cfg overrides: KEY VALUE ...")
    return parser


def load_detections(output_dir):
    path = os.path.join(output_dir, "detections.pkl")
    with open(path, "rb") as f:
        return pickle.load(f)


def rescore(imdb_name, output_dir, comp_mode=False, matlab_eval=False, nms=False):
    """all_boxes from <output_dir>/detections.pkl -> imdb.evaluate_detections; returns what the imdb returns."""
    dataset = get_imdb(imdb_name)
    dataset.competition_mode(comp_mode)
    dataset.config["matlab_eval"] = matlab_eval
    all_boxes = load_detections(output_dir)
    if nms:
        from model.test import apply_nms
        print("Applying NMS to all detections")
        all_boxes = apply_nms(all_boxes, cfg.TEST.NMS)
    print("Evaluating detections")
    return dataset.evaluate_detections(all_boxes, output_dir)


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        parser.print_help()
        return 1
    args = parser.parse_args(argv)
    if args.set_cfgs:
        cfg_from_list(args.set_cfgs)
    rescore(args.imdb_name, os.path.abspath(args.output_dir[0]), args.comp_mode, args.matlab_eval, args.apply_nms)
    return 0


if __name__ == "__main__":
    sys.exit(main())
