"""Box bookkeeping helpers of the dataset layer (lib/datasets/ds_utils.py:13-49), same names and results.
Host-side index / format utilities for ground truth and result files -- nothing here is on the detection path."""
import numpy as np


def unique_boxes(boxes, scale=1.0):
    """Indices of one representative per distinct (rounded, scaled) box, in ascending order."""
    keys = np.round(np.asarray(boxes) * scale).dot(np.array([1, 1e3, 1e6, 1e9]))
    return np.sort(np.unique(keys, return_index=True)[1])


def xywh_to_xyxy(boxes):
    """[x, y, w, h] -> [x1, y1, x2, y2] with inclusive pixel corners (x2 = x + w - 1)."""
    boxes = np.asarray(boxes)
    return np.hstack((boxes[:, 0:2], boxes[:, 0:2] + boxes[:, 2:4] - 1))


def xyxy_to_xywh(boxes):
    """Inverse of xywh_to_xyxy (w = x2 - x1 + 1)."""
    boxes = np.asarray(boxes)
    return np.hstack((boxes[:, 0:2], boxes[:, 2:4] - boxes[:, 0:2] + 1))


def validate_boxes(boxes, width=0, height=0):
    """Asserts 0 <= x1 <= x2 < width and 0 <= y1 <= y2 < height for every row."""
    boxes = np.asarray(boxes)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    assert (x1 >= 0).all() and (y1 >= 0).all()
    assert (x2 >= x1).all() and (y2 >= y1).all()
    assert (x2 < width).all() and (y2 < height).all()


def filter_small_boxes(boxes, min_size):
    """Indices of boxes at least min_size wide and high (exclusive extents, as the reference)."""
    boxes = np.asarray(boxes)
    return np.where((boxes[:, 2] - boxes[:, 0] >= min_size) & (boxes[:, 3] - boxes[:, 1] > min_size))[0]
