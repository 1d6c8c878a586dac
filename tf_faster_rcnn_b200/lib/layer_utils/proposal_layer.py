"""`proposal_layer` / `proposal_layer_tf` with the reference's signatures (lib/layer_utils/proposal_layer.py:16-52, :55-83), for
callers that hold the RPN outputs as arrays.  The network itself never calls these: its graph runs frcnn_rpn_decode ->
frcnn_sort_desc -> frcnn_proposals straight from the RPN head's device buffer (engine.ShapePlan).  Here the decode + clip are
the host functions of model.bbox_transform, while the score sort and the greedy NMS -- the expensive part -- run in the same
device kernels as the graph's (tf_faster_rcnn_b200/csrc/sort.cu, nms.cu): there is no CPU NMS in this build.

Returns (blob [k,5] = (0, x1,y1,x2,y2) fp32, scores [k,1]) as NumPy arrays."""
import numpy as np
import torch

from model.config import cfg
from model.bbox_transform import bbox_transform_inv, clip_boxes
from tf_faster_rcnn_b200 import engine, ops, _native as N


def _select(proposals, scores, pre_nms_top_n, post_nms_top_n, thresh, flags):
    """Device part: stable descending sort of `scores`, greedy NMS over the first pre_nms_top_n, first post_nms_top_n survivors."""
    n = int(scores.shape[0])
    post = int(post_nms_top_n) if post_nms_top_n > 0 else n
    if n == 0:
        return np.zeros((0, 5), np.float32), np.zeros((0, 1), np.float32)
    if thresh >= 0 and post > 1024:
        raise ValueError("post-NMS top-N %d exceeds the proposal kernel's capacity (1024)" % post)
    pd = torch.from_numpy(np.ascontiguousarray(proposals, dtype=np.float32)).cuda()
    sd = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32).ravel()).cuda()
    order = torch.empty(n, dtype=torch.int32, device="cuda"); sk = torch.empty(n, dtype=torch.float32, device="cuda")
    ops.sort_desc(sd, order, sk)
    rois = torch.empty((post, 5), dtype=torch.float32, device="cuda"); rs = torch.empty(post, dtype=torch.float32, device="cuda")
    keep = torch.empty(post, dtype=torch.int32, device="cuda"); num = torch.empty(1, dtype=torch.int32, device="cuda")
    ops.proposals(pd, sd, order, int(pre_nms_top_n), post, thresh, flags, rois, rs, keep, num)
    k = int(num.item())
    return rois[:k].cpu().numpy(), rs[:k].cpu().numpy().reshape(-1, 1)


def _decode(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, num_anchors):
    scores = np.asarray(rpn_cls_prob)[:, :, :, num_anchors:].reshape(-1)
    deltas = np.asarray(rpn_bbox_pred, dtype=np.float32).reshape(-1, 4)
    proposals = clip_boxes(bbox_transform_inv(np.asarray(anchors, dtype=np.float32), deltas), im_info[:2])
    return proposals, scores


def proposal_layer(rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key, _feat_stride, anchors, num_anchors):
    """NumPy-path semantics: top RPN_PRE_NMS_TOP_N by score, '+1' NMS with the cpu_nms / gpu_nms predicate (cfg.USE_GPU_NMS),
    first RPN_POST_NMS_TOP_N survivors."""
    if isinstance(cfg_key, bytes):
        cfg_key = cfg_key.decode("utf-8")
    c = cfg[cfg_key]
    proposals, scores = _decode(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, num_anchors)
    thr, flags = engine.nms_threshold(c.RPN_NMS_THRESH, bool(cfg.USE_GPU_NMS))
    return _select(proposals, scores, c.RPN_PRE_NMS_TOP_N, c.RPN_POST_NMS_TOP_N, thr, flags)


def proposal_layer_tf(rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key, _feat_stride, anchors, num_anchors):
    """tf.image.non_max_suppression semantics over ALL anchors (no pre-NMS cut, continuous areas, strict >, degenerate boxes kept)."""
    if isinstance(cfg_key, bytes):
        cfg_key = cfg_key.decode("utf-8")
    c = cfg[cfg_key]
    proposals, scores = _decode(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, num_anchors)
    return _select(proposals, scores, 0, c.RPN_POST_NMS_TOP_N, float(np.float32(c.RPN_NMS_THRESH)), N.NMS_MODE_TF)
