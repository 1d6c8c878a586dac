"""Anchor tiling over the feature grid (lib/layer_utils/snippets.py:14-49).

On the device path anchors are never materialised: frcnn_rpn_decode adds (x*16, y*16) to the base
anchor inside the kernel.  This host function exists for callers that want the explicit table."""
import numpy as np

from layer_utils.generate_anchors import generate_anchors


def generate_anchors_pre(height, width, feat_stride, anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2)):
    base = generate_anchors(ratios=np.array(anchor_ratios), scales=np.array(anchor_scales))
    xs = np.arange(width) * feat_stride
    ys = np.arange(height) * feat_stride
    gx, gy = np.meshgrid(xs, ys)
    shift = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)
    table = (shift[:, None, :] + base[None, :, :]).reshape(-1, 4).astype(np.float32)
    return table, np.int32(table.shape[0])
