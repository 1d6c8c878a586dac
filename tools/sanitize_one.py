"""One small ResNet-50 image (and a batch of 2) through the whole device path, eager launches (no CUDA graph) -- the workload for
`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_one.py` (profiles/r02_sanitizer.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tf_faster_rcnn_b200 import paths, synth  # noqa: E402

paths.add_lib_path()
from model.config import cfg  # noqa: E402
from nets.resnet_v1 import resnetv1  # noqa: E402


def main():
    hw = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 160)
    cfg.TEST.HAS_RPN = True
    net = resnetv1(num_layers=50)
    net.use_cuda_graph = False
    net.create_architecture("TEST", 21, tag="default", anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2))
    net.load_weights(synth.make("res50", 21, 9))
    blob = synth.synthetic_blob(*hw)
    det, _ = net.detect(blob, np.array([hw[0], hw[1], 1.0], np.float32), hw)
    dets, _ = net.detect_batch(np.concatenate([blob, blob], axis=0), [1.0, 1.0], [hw, hw])
    print("sanitize_one: %dx%d -> %d detections; batch of 2 -> %s" % (hw[0], hw[1], det.shape[0], [d.shape[0] for d in dets]))


if __name__ == "__main__":
    main()
