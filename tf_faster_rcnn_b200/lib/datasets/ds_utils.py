"""Box bookkeeping helpers of the dataset layer, same names and results as the reference's lib/datasets/ds_utils.py:13-49
(checked against it in tests/test_datasets.py).  Host-side index / format utilities for ground truth and result files;
nothing here is on the detection path."""
import numpy as np

# unique_boxes folds a rounded box into one float64 key, one decimal "digit group" of 1000 per coordinate -- including the
# reference's behaviour that coordinates >= 1000 can collide.
_KEY_WEIGHTS = np.array([1.0, 1.0e3, 1.0e6, 1.0e9])


def _as_boxes(boxes):
    arr = np.asarray(boxes)
    return arr.reshape(-1, 4) if arr.ndim != 2 else arr


def unique_boxes(boxes, scale=1.0):
    """Ascending indices of the first occurrence of every distinct box after `round(box * scale)`."""
    keys = np.round(_as_boxes(boxes) * scale) @ _KEY_WEIGHTS
    _, first = np.unique(keys, return_index=True)
    first.sort()
    return first


def xywh_to_xyxy(boxes):
    """[x, y, w, h] rows -> [x1, y1, x2, y2] with inclusive corners: x2 = x + w - 1, y2 = y + h - 1."""
    b = _as_boxes(boxes)
    corner, extent = b[:, :2], b[:, 2:4]
    return np.concatenate([corner, corner + extent - 1], axis=1)


def xyxy_to_xywh(boxes):
    """Inverse of xywh_to_xyxy: w = x2 - x1 + 1, h = y2 - y1 + 1."""
    b = _as_boxes(boxes)
    lo, hi = b[:, :2], b[:, 2:4]
    return np.concatenate([lo, hi - lo + 1], axis=1)


def validate_boxes(boxes, width=0, height=0):
    """AssertionError unless every row satisfies 0 <= x1 <= x2 < width and 0 <= y1 <= y2 < height."""
    b = _as_boxes(boxes)
    for lo, hi, limit in ((b[:, 0], b[:, 2], width), (b[:, 1], b[:, 3], height)):
        assert np.all(lo >= 0), "negative coordinate"
        assert np.all(hi >= lo), "inverted box"
        assert np.all(hi < limit), "box outside the image"


def filter_small_boxes(boxes, min_size):
    """Indices of rows with x2 - x1 >= min_size and y2 - y1 > min_size (the asymmetry is the reference's)."""
    b = _as_boxes(boxes)
    wide_enough = (b[:, 2] - b[:, 0]) >= min_size
    tall_enough = (b[:, 3] - b[:, 1]) > min_size
    return np.flatnonzero(wide_enough & tall_enough)
