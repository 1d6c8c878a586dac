"""Oracle: box codec in fp32 with separate mul/add roundings (test infrastructure).

Follows lib/model/bbox_transform.py:35-65 (decode), :68-81 / :110-115 (two-sided
clip used by the RPN stage) and lib/model/test.py:67-77 (one-sided clip used
after the head).  numpy evaluates every binary op with its own fp32 rounding
(no FMA); the CUDA kernels reproduce exactly that op order.
"""
import numpy as np

F = np.float32


def exp_f32(x):
    """fp32 exp, correctly rounded (computed in fp64, rounded once).  The reference calls np.exp /
    tf.exp on fp32 (bbox_transform.py:52-53,103-104), each a <=1-ulp libm-grade exp that differs between
    numpy/Eigen builds; the oracle pins the one well-defined member of that family."""
    return np.exp(np.asarray(x, dtype=np.float64)).astype(F)


def decode(boxes, deltas):
    """bbox_transform_inv: boxes fp32 [R,4], deltas fp32 [R,4K] -> fp32 [R,4K]."""
    boxes = np.ascontiguousarray(boxes, dtype=F)
    deltas = np.ascontiguousarray(deltas, dtype=F)
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=F)
    w = (boxes[:, 2] - boxes[:, 0] + F(1.0))[:, None]
    h = (boxes[:, 3] - boxes[:, 1] + F(1.0))[:, None]
    cx = boxes[:, 0:1] + F(0.5) * w
    cy = boxes[:, 1:2] + F(0.5) * h
    pcx = deltas[:, 0::4] * w + cx
    pcy = deltas[:, 1::4] * h + cy
    pw = exp_f32(deltas[:, 2::4]) * w
    ph = exp_f32(deltas[:, 3::4]) * h
    out = np.empty_like(deltas)
    out[:, 0::4] = pcx - F(0.5) * pw
    out[:, 1::4] = pcy - F(0.5) * ph
    out[:, 2::4] = pcx + F(0.5) * pw
    out[:, 3::4] = pcy + F(0.5) * ph
    return out


def clip_two_sided(boxes, im_h, im_w):
    """clip_boxes / clip_boxes_tf: max(min(v, dim-1), 0) on every coordinate."""
    boxes = boxes.copy()
    xmax = F(im_w) - F(1.0)
    ymax = F(im_h) - F(1.0)
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], xmax), F(0))
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], ymax), F(0))
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], xmax), F(0))
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], ymax), F(0))
    return boxes


def clip_one_sided(boxes, im_h, im_w):
    """test.py:_clip_boxes: x1,y1 >= 0 only; x2,y2 <= dim-1 only (ORIGINAL image dims, ints)."""
    boxes = boxes.copy()
    boxes[:, 0::4] = np.maximum(boxes[:, 0::4], F(0))
    boxes[:, 1::4] = np.maximum(boxes[:, 1::4], F(0))
    boxes[:, 2::4] = np.minimum(boxes[:, 2::4], F(im_w - 1))
    boxes[:, 3::4] = np.minimum(boxes[:, 3::4], F(im_h - 1))
    return boxes
