"""`proposal_top_layer` / `proposal_top_layer_tf` (lib/layer_utils/proposal_top_layer.py:17-55, :58-85): the TEST.MODE='top'
selection -- the cfg.TEST.RPN_TOP_N best-scoring anchors, no NMS -- for callers holding the RPN outputs as arrays (the network's
own graph does this with frcnn_sort_desc + frcnn_proposals(thresh < 0)).  The score sort runs on the device; the decode of the
selected anchors is model.bbox_transform's host code."""
import numpy as np
import numpy.random as npr
import torch

from model.config import cfg
from model.bbox_transform import bbox_transform_inv, clip_boxes
from tf_faster_rcnn_b200 import ops


def _top_indices(scores, k):
    n = int(scores.shape[0])
    sd = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).cuda()
    order = torch.empty(n, dtype=torch.int32, device="cuda"); sk = torch.empty(n, dtype=torch.float32, device="cuda")
    ops.sort_desc(sd, order, sk)
    return order[:k].cpu().numpy().astype(np.int64)


def proposal_top_layer(rpn_cls_prob, rpn_bbox_pred, im_info, _feat_stride, anchors, num_anchors):
    top_n = int(cfg.TEST.RPN_TOP_N)
    scores = np.asarray(rpn_cls_prob)[:, :, :, num_anchors:].reshape(-1)
    deltas = np.asarray(rpn_bbox_pred, dtype=np.float32).reshape(-1, 4)
    if scores.shape[0] < top_n:
        # fewer anchors than requested: the reference samples WITH replacement (proposal_top_layer.py:31-34)
        top = npr.choice(scores.shape[0], size=top_n, replace=True)
    else:
        top = _top_indices(scores, top_n)
    proposals = clip_boxes(bbox_transform_inv(np.asarray(anchors, dtype=np.float32)[top], deltas[top]), im_info[:2])
    blob = np.hstack((np.zeros((proposals.shape[0], 1), np.float32), proposals.astype(np.float32, copy=False)))
    return blob, scores[top].reshape(-1, 1)


def proposal_top_layer_tf(rpn_cls_prob, rpn_bbox_pred, im_info, _feat_stride, anchors, num_anchors):
    """tf.nn.top_k needs k <= number of anchors (it raises otherwise, like TensorFlow)."""
    top_n = int(cfg.TEST.RPN_TOP_N)
    scores = np.asarray(rpn_cls_prob)[:, :, :, num_anchors:].reshape(-1)
    if scores.shape[0] < top_n:
        raise ValueError("top_k: k=%d exceeds the %d anchors" % (top_n, scores.shape[0]))
    deltas = np.asarray(rpn_bbox_pred, dtype=np.float32).reshape(-1, 4)
    top = _top_indices(scores, top_n)
    proposals = clip_boxes(bbox_transform_inv(np.asarray(anchors, dtype=np.float32)[top], deltas[top]), im_info[:2])
    blob = np.hstack((np.zeros((top_n, 1), np.float32), proposals.astype(np.float32, copy=False)))
    return blob, scores[top].reshape(-1, 1)
