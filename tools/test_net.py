#!/usr/bin/env python
"""Run the detector over an imdb (counterpart of the reference's tools/test_net.py:58-122; same flags).

    python tools/test_net.py --imdb synthetic_8_21 --net res101 [--cfg x.yml] [--model ckpt] [--set K V ...]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/test_net.py --imdb voc_2007_test --net res101 --model ckpt
        (one process per GPU: images are sharded i mod W, records all-gathered per step, rank 0 evaluates)"""
import argparse
import os
import pprint
import sys

import _init_paths  # noqa: F401
import numpy as np

from model.config import cfg, cfg_from_file, cfg_from_list
from model.test import test_net
from datasets.factory import get_imdb
from tools_common import build_net  # noqa: E402


def parse_args():
    p = argparse.ArgumentParser(description='Test a Faster R-CNN network on the B200 path')
    p.add_argument('--cfg', dest='cfg_file', default=None, type=str)
    p.add_argument('--model', dest='model', default=None, type=str)
    p.add_argument('--imdb', dest='imdb_name', default='synthetic_4_21', type=str)
    p.add_argument('--comp', dest='comp_mode', action='store_true')
    p.add_argument('--num_dets', dest='max_per_image', default=100, type=int)
    p.add_argument('--tag', dest='tag', default='', type=str)
    p.add_argument('--net', dest='net', default='res50', type=str, help='vgg16, res50, res101, res152, mobile')
    p.add_argument('--batch', dest='batch', default=1, type=int,
                   help='images per device launch: consecutive images with equal blob shapes are grouped (not in the reference)')
    p.add_argument('--set', dest='set_cfgs', default=None, nargs=argparse.REMAINDER)
    return p.parse_args()


def init_distributed():
    """One process per GPU when launched by torchrun (WORLD_SIZE > 1); a plain launch stays single-process."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    return dist.get_rank()


if __name__ == '__main__':
    args = parse_args()
    rank = init_distributed()
    if args.cfg_file is not None:
        cfg_from_file(args.cfg_file)
    if args.set_cfgs is not None:
        cfg_from_list(args.set_cfgs)
    if rank == 0:
        pprint.pprint({k: cfg[k] for k in ("TEST", "ANCHOR_SCALES", "ANCHOR_RATIOS", "USE_GPU_NMS", "USE_E2E_TF")})
    filename = os.path.splitext(os.path.basename(args.model))[0] if args.model else 'default'
    tag = args.tag if args.tag else 'default'
    imdb = get_imdb(args.imdb_name)
    imdb.competition_mode(args.comp_mode)
    net = build_net(args.net, imdb.num_classes, args.model)
    import model.test as model_test
    model_test.BATCH_SIZE = max(1, args.batch)
    test_net(None, net, imdb, filename + '/' + tag, max_per_image=args.max_per_image)
