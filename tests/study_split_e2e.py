#!/usr/bin/env python
"""CPU study (not a test; run by hand): end-to-end effect of the GEMM operand-split scheme on the detector's outputs.
The oracle's conv / FC primitives are swapped for float64 evaluations of tensor-core-rounded operands
(tools/numerics_split_schemes.py explains the schemes), the whole TEST graph runs per scheme, and every variant is
compared with the 'exact' run (float64 products of the unrounded fp32 operands, rounded to fp32 once per layer).

    python tests/study_split_e2e.py [res50|vgg16] [H W]

Lives under tests/ because it imports the oracle (test infrastructure)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import layers as L, pipeline as P  # noqa: E402
from tf_faster_rcnn_b200 import synth  # noqa: E402
from numerics_split_schemes import tf32, bf16  # noqa: E402

_conv_fp32, _fc_fp32 = L.conv2d, L.fully_connected


def _terms(scheme, a, b):
    """[(A, B)] operand pairs whose exact products sum to the emulated result."""
    if scheme == "exact":
        return [(a, b)]
    ah, bh = tf32(a), tf32(b)
    if scheme == "1xtf32":
        return [(ah, bh)]
    if scheme == "3xtf32":
        return [(ah, bh), (tf32(a - ah), bh), (ah, tf32(b - bh))]
    if scheme == "tf32+2xbf16":
        return [(ah, bh), (bf16(a - ah), bf16(bh)), (bf16(ah), bf16(b - bh))]
    raise ValueError(scheme)


def install(scheme):
    if scheme == "fp32":
        L.conv2d, L.fully_connected = _conv_fp32, _fc_fp32
        return

    def conv2d(x, w_hwio, stride=1, padding="SAME", groups=1):
        if groups != 1:
            return _conv_fp32(x, w_hwio, stride, padding, groups)
        kh, kw = w_hwio.shape[:2]
        pads = (0, 0, 0, 0)
        if padding == "SAME":
            pt, pb = L.same_pads(x.shape[1], kh, stride)
            pl, pr = L.same_pads(x.shape[2], kw, stride)
            pads = (pl, pr, pt, pb)
        out = None
        for a, b in _terms(scheme, np.asarray(x, np.float32), np.asarray(w_hwio, np.float32)):
            xt = TF.pad(torch.from_numpy(np.ascontiguousarray(a)).permute(0, 3, 1, 2).double(), pads)
            wt = torch.from_numpy(np.ascontiguousarray(b)).permute(3, 2, 0, 1).double().contiguous()
            y = TF.conv2d(xt, wt, None, stride=stride)
            out = y if out is None else out + y
        return np.ascontiguousarray(out.permute(0, 2, 3, 1).float().numpy())

    def fully_connected(x, w_io):
        out = None
        for a, b in _terms(scheme, np.asarray(x, np.float32), np.asarray(w_io, np.float32)):
            y = torch.from_numpy(np.ascontiguousarray(a)).double() @ torch.from_numpy(np.ascontiguousarray(b)).double()
            out = y if out is None else out + y
        return out.float().numpy()

    L.conv2d, L.fully_connected = conv2d, fully_connected


def main():
    net = sys.argv[1] if len(sys.argv) > 1 else "res50"
    hw = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (192, 256)
    C, scales = 21, (8, 16, 32)
    w = synth.make(net, C, 3 * len(scales))
    blob = synth.synthetic_blob(*hw)
    im_info = np.array([hw[0], hw[1], 1.0], np.float32)
    runs = {}
    for scheme in ("exact", "fp32", "3xtf32", "tf32+2xbf16", "1xtf32"):
        install(scheme)
        runs[scheme] = P.test_image(net, w, blob, im_info, C, P.opts(anchor_scales=scales))
        print("ran", scheme, flush=True)
    install("fp32")
    ref = runs["exact"]
    print("%s %dx%d, %d classes; every column vs the exact run" % (net, hw[0], hw[1], C))
    print("%-12s %10s %10s %12s %12s %12s" % ("scheme", "rois same", "order same", "feat relerr", "cls_prob err", "bbox_pred err"))
    for scheme, st in runs.items():
        common, ia, ib = np.intersect1d(st["roi_keep"], ref["roi_keep"], return_indices=True)
        same_order = bool(len(st["roi_keep"]) == len(ref["roi_keep"]) and np.array_equal(st["roi_keep"], ref["roi_keep"]))
        feat = float(np.abs(st["feat"] - ref["feat"]).max() / np.abs(ref["feat"]).max())
        prob = float(np.abs(st["cls_prob"][ia] - ref["cls_prob"][ib]).max())
        box = float(np.abs(st["bbox_pred"][ia] - ref["bbox_pred"][ib]).max())
        print("%-12s %6d/%-3d %10s %12.2e %12.2e %12.2e" % (scheme, len(common), len(ref["roi_keep"]), same_order, feat, prob, box))


if __name__ == "__main__":
    main()
