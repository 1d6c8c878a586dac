"""Base anchor windows (host-side constants uploaded once per network).

Closed form of lib/layer_utils/generate_anchors.py:41-105: the reference window is (0,0,15,15)
(area 256, centre 7.5); for every ratio r the window becomes ws x hs with ws = round(sqrt(256/r)),
hs = round(ws*r) (np.round: half to even), and every scale s multiplies both; anchors are centred
on 7.5.  Row order: ratios major, scales minor.  Known answer: generate_anchors.py:14-39 (MATLAB, 1-based).
"""
import numpy as np


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    ctr = 0.5 * (base_size - 1)
    ws = np.round(np.sqrt(float(base_size * base_size) / ratios))
    hs = np.round(ws * ratios)
    w = (ws[:, None] * scales[None, :]).ravel()
    h = (hs[:, None] * scales[None, :]).ravel()
    return np.stack([ctr - 0.5 * (w - 1), ctr - 0.5 * (h - 1), ctr + 0.5 * (w - 1), ctr + 0.5 * (h - 1)], axis=1)
