"""Box codec with the reference's function names (lib/model/bbox_transform.py): bbox_transform :14-32 (regression targets),
bbox_transform_inv :35-65 / bbox_transform_inv_tf :84-107 (decode), clip_boxes :68-81 / clip_boxes_tf :110-115 (two-sided clip).

On the inference path the decode + clip run inside device kernels (frcnn_rpn_decode for the RPN, frcnn_bbox_decode for the
im_detect tail); these host functions are the module surface other code imports (layer_utils.proposal_layer, user scripts)
and work on NumPy arrays in the arrays' own dtype, one rounding per operation like NumPy itself.  The `_tf` names take and
return NumPy arrays too: there is no TensorFlow graph in this build."""
import numpy as np


def _centre_size(boxes):
    w = boxes[:, 2] - boxes[:, 0] + 1.0
    h = boxes[:, 3] - boxes[:, 1] + 1.0
    return boxes[:, 0] + 0.5 * w, boxes[:, 1] + 0.5 * h, w, h


def bbox_transform(ex_rois, gt_rois):
    """Regression targets (dx, dy, dw, dh) that map ex_rois onto gt_rois; rows of [N,4]."""
    ecx, ecy, ew, eh = _centre_size(ex_rois)
    gcx, gcy, gw, gh = _centre_size(gt_rois)
    return np.stack(((gcx - ecx) / ew, (gcy - ecy) / eh, np.log(gw / ew), np.log(gh / eh)), axis=1)


def bbox_transform_inv(boxes, deltas):
    """boxes [N,4], deltas [N,4K] -> predicted boxes [N,4K] (x1,y1,x2,y2 per class)."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    cx, cy, w, h = (v[:, np.newaxis] for v in _centre_size(boxes))
    pcx = deltas[:, 0::4] * w + cx
    pcy = deltas[:, 1::4] * h + cy
    pw = np.exp(deltas[:, 2::4]) * w
    ph = np.exp(deltas[:, 3::4]) * h
    out = np.zeros(deltas.shape, dtype=deltas.dtype)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw
    out[:, 3::4] = pcy + 0.5 * ph
    return out


def clip_boxes(boxes, im_shape):
    """Clamp every coordinate into the image: x in [0, im_shape[1]-1], y in [0, im_shape[0]-1] (in place, like the reference)."""
    xmax, ymax = im_shape[1] - 1, im_shape[0] - 1
    for first, hi in ((0, xmax), (1, ymax), (2, xmax), (3, ymax)):
        boxes[:, first::4] = np.maximum(np.minimum(boxes[:, first::4], hi), 0)
    return boxes


def bbox_transform_inv_tf(boxes, deltas):
    return bbox_transform_inv(np.asarray(boxes, dtype=np.float32), np.asarray(deltas, dtype=np.float32))


def clip_boxes_tf(boxes, im_info):
    return clip_boxes(np.array(boxes, dtype=np.float32), im_info)
