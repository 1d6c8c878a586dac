"""Oracle: anchor enumeration (test infrastructure, see oracle/__init__.py).

Follows lib/layer_utils/generate_anchors.py:41-105 (base anchors) and
lib/layer_utils/snippets.py:14-49 (tiling over the h x w grid, (h, w, a) order).
"""
import numpy as np


def _centre_form(box):
    """(x1,y1,x2,y2) -> (w, h, cx, cy) with the +1 pixel convention (generate_anchors.py:52-61)."""
    w = box[2] - box[0] + 1.0
    h = box[3] - box[1] + 1.0
    return w, h, box[0] + 0.5 * (w - 1.0), box[1] + 0.5 * (h - 1.0)


def _corner_form(ws, hs, cx, cy):
    """generate_anchors.py:64-76."""
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    return np.hstack([cx - 0.5 * (ws - 1), cy - 0.5 * (hs - 1),
                      cx + 0.5 * (ws - 1), cy + 0.5 * (hs - 1)])


def base_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    """ratios-major x scales base anchors, float64, np.round = half-to-even (generate_anchors.py:79-105)."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w, h, cx, cy = _centre_form(np.array([0.0, 0.0, base_size - 1.0, base_size - 1.0]))
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    per_ratio = _corner_form(ws, hs, cx, cy)
    out = []
    for row in per_ratio:
        w, h, cx, cy = _centre_form(row)
        out.append(_corner_form(w * scales, h * scales, cx, cy))
    return np.vstack(out)


def tiled_anchors(height, width, feat_stride=16, scales=(8, 16, 32), ratios=(0.5, 1, 2)):
    """All anchors as fp32 [height*width*A, 4], location-major / anchor-minor (snippets.py:14-49).

    The TF path builds them as int32 and casts (snippets.py:44-49); values are integral so
    the float64 -> float32 path of the numpy variant gives identical numbers.
    """
    base = base_anchors(ratios=ratios, scales=scales)
    sx, sy = np.meshgrid(np.arange(width) * feat_stride, np.arange(height) * feat_stride)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], axis=1)
    allb = base[None, :, :] + shifts[:, None, :]
    return allb.reshape(-1, 4).astype(np.float32)
