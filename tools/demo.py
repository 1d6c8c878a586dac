#!/usr/bin/env python
"""Detection demo on a directory of images (counterpart of the reference's tools/demo.py, whose own script also
runs unchanged against this repo's lib/: see INTEGRATION.md).  Prints the detections above --conf per class.

    python tools/demo.py --net res101 --images <dir> [--model <ckpt prefix>] [--classes 21]"""
import argparse
import os

import _init_paths  # noqa: F401
import cv2
import numpy as np

from model.config import cfg
from model.test import im_detect
from model.nms_wrapper import nms
from utils.timer import Timer
from nets.vgg16 import vgg16
from nets.resnet_v1 import resnetv1
from nets.mobilenet_v1 import mobilenetv1
from tf_faster_rcnn_b200 import checkpoint, synth


def build(net_name, num_classes, model=None):
    cfg.TEST.HAS_RPN = True
    net = vgg16() if net_name == "vgg16" else mobilenetv1() if net_name == "mobile" else resnetv1(int(net_name[3:]))
    net.create_architecture("TEST", num_classes, tag="default", anchor_scales=cfg.ANCHOR_SCALES, anchor_ratios=cfg.ANCHOR_RATIOS)
    if model:
        net.load_weights(checkpoint.load_variables(model), strict=True)   # TF V2 bundle or .npz
    else:
        net.load_weights(synth.make(net_name, num_classes, net.num_anchors))
    return net


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="res101")
    ap.add_argument("--images", required=True)
    ap.add_argument("--model", default=None)
    ap.add_argument("--classes", type=int, default=21)
    ap.add_argument("--conf", type=float, default=0.8)
    ap.add_argument("--nms", type=float, default=0.3)
    a = ap.parse_args()
    net = build(a.net, a.classes, a.model)
    for f in sorted(os.listdir(a.images)):
        if not f.lower().endswith((".jpg", ".jpeg", ".png")):
            continue
        im = cv2.imread(os.path.join(a.images, f))
        t = Timer(); t.tic()
        scores, boxes = im_detect(None, net, im)
        t.toc()
        print("Detection took {:.3f}s for {:d} object proposals ({})".format(t.total_time, boxes.shape[0], f))
        for c in range(1, a.classes):
            dets = np.hstack((boxes[:, 4 * c:4 * c + 4], scores[:, c][:, None])).astype(np.float32)
            dets = dets[nms(dets, a.nms), :]
            for d in dets[dets[:, 4] >= a.conf]:
                print("  class %d  %.3f  [%.1f %.1f %.1f %.1f]" % (c, d[4], d[0], d[1], d[2], d[3]))
