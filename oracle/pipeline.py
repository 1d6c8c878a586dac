"""Oracle: the whole TEST-mode path, stage by stage (test infrastructure).

  rpn_head              lib/nets/network.py:323-359 (+ _reshape_layer :68-78, _softmax_layer :80-86)
  rpn_decode            lib/layer_utils/proposal_layer.py:62-69 / :26-31
  proposals_e2e_tf      lib/layer_utils/proposal_layer.py:56-83   (USE_E2E_TF=True, the default)
  proposals_numpy       lib/layer_utils/proposal_layer.py:16-53   (USE_E2E_TF=False)
  proposals_top         lib/layer_utils/proposal_top_layer.py:58-85 (TEST.MODE='top', TF variant)
  crop_pool             lib/nets/network.py:141-157, lib/nets/resnet_v1.py:55-76
  region_classification lib/nets/network.py:361-378, :428-432
  test_image            lib/nets/network.py:233-262, :470-479
  get_image_blob        lib/model/test.py:26-58, lib/utils/blob.py:17-30
  im_detect_post        lib/model/test.py:95-107
  test_net_post         lib/model/test.py:162-180
"""
import numpy as np
from . import layers as L
from . import nets as N
from . import anchors as A
from . import boxes as B
from . import nms as NMS

F = np.float32

DEFAULTS = dict(
    rpn_nms_thresh=0.7, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, rpn_top_n=5000,
    test_mode="nms", use_e2e_tf=True, use_gpu_nms=False, pooling_size=7, resnet_max_pool=False,
    bbox_stds=(0.1, 0.1, 0.2, 0.2), bbox_means=(0.0, 0.0, 0.0, 0.0),
    anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2), nms_thresh=0.3, max_per_image=100,
    pixel_means=(102.9801, 115.9465, 122.7717), test_scale=600, test_max_size=1000,
)


def opts(**kw):
    o = dict(DEFAULTS)
    o.update(kw)
    return o


def rpn_head(net, w, feat):
    sc = N.scope_of(net)
    rpn = L.relu(L.bias_add(L.conv2d(feat, w[sc + "/rpn_conv/3x3/weights"], 1, "SAME"), w[sc + "/rpn_conv/3x3/biases"]))
    cls = L.bias_add(L.conv2d(rpn, w[sc + "/rpn_cls_score/weights"], 1, "VALID"), w[sc + "/rpn_cls_score/biases"])
    box = L.bias_add(L.conv2d(rpn, w[sc + "/rpn_bbox_pred/weights"], 1, "VALID"), w[sc + "/rpn_bbox_pred/biases"])
    return rpn, cls, box


def rpn_fg_prob(rpn_cls_score):
    """[1,h,w,2A] -> fg probability [h*w*A] in (h,w,a) order: softmax over the pair
    (channel a = bg, channel A+a = fg) -- what the reshape/softmax/reshape dance computes."""
    a = rpn_cls_score.shape[3] // 2
    pair = np.stack([rpn_cls_score[0, :, :, :a], rpn_cls_score[0, :, :, a:]], axis=-1)
    return L.softmax_lastdim(pair)[..., 1].reshape(-1).astype(F)


def rpn_decode(rpn_cls_score, rpn_bbox_pred, im_info, o):
    """-> scores [N], proposals [N,4] (decoded + clipped to the blob), anchors [N,4]."""
    _, h, w, _ = rpn_cls_score.shape
    anc = A.tiled_anchors(h, w, 16, o["anchor_scales"], o["anchor_ratios"])
    scores = rpn_fg_prob(rpn_cls_score)
    deltas = rpn_bbox_pred.reshape(-1, 4)
    props = B.clip_two_sided(B.decode(anc, deltas), im_info[0], im_info[1])
    return scores, props, anc


def proposals_e2e_tf(scores, props, o):
    keep = NMS.nms_tf_c(props, scores, o["rpn_post_nms_top_n"], o["rpn_nms_thresh"])
    rois = np.hstack([np.zeros((keep.shape[0], 1), F), props[keep]]).astype(F)
    return rois, scores[keep].reshape(-1, 1), keep


def proposals_numpy(scores, props, o):
    order = NMS.argsort_desc(scores)
    if o["rpn_pre_nms_top_n"] > 0:
        order = order[:o["rpn_pre_nms_top_n"]]
    p, s = props[order], scores[order]
    keep = NMS.nms_plus1_c(np.hstack([p, s[:, None]]), o["rpn_nms_thresh"], inclusive=not o["use_gpu_nms"])
    if o["rpn_post_nms_top_n"] > 0:
        keep = keep[:o["rpn_post_nms_top_n"]]
    rois = np.hstack([np.zeros((keep.shape[0], 1), F), p[keep]]).astype(F)
    return rois, s[keep].reshape(-1, 1), order[keep]


def proposals_top(scores, props, o):
    """TF variant: top_k (ties -> lower index), gather, no NMS. (decode/clip commute with gather.)"""
    keep = NMS.argsort_desc(scores)[:o["rpn_top_n"]]
    rois = np.hstack([np.zeros((keep.shape[0], 1), F), props[keep]]).astype(F)
    return rois, scores[keep].reshape(-1, 1), keep


def proposals(scores, props, o):
    if o["test_mode"] == "top":
        return proposals_top(scores, props, o)
    return proposals_e2e_tf(scores, props, o) if o["use_e2e_tf"] else proposals_numpy(scores, props, o)


def roi_norm_boxes(feat_shape, rois):
    """network.py:146-153: normalised [y1,x1,y2,x2] for crop_and_resize."""
    hh = (F(feat_shape[1]) - F(1.0)) * F(16.0)
    ww = (F(feat_shape[2]) - F(1.0)) * F(16.0)
    return np.stack([rois[:, 2] / hh, rois[:, 1] / ww, rois[:, 4] / hh, rois[:, 3] / ww], axis=1).astype(F)


def crop_pool(net, feat, rois, o):
    nb = roi_norm_boxes(feat.shape, rois)
    p = o["pooling_size"]
    if net.startswith("res") and not o["resnet_max_pool"]:
        return L.crop_and_resize(feat, nb, p)
    return L.max_pool(L.crop_and_resize(feat, nb, 2 * p), 2, 2, "SAME")


def region_classification(net, w, fc7, num_classes, o):
    sc = N.scope_of(net)
    cls_score = L.fully_connected(fc7, w[sc + "/cls_score/weights"]) + w[sc + "/cls_score/biases"]
    cls_prob = L.softmax_lastdim(cls_score)
    bbox = L.fully_connected(fc7, w[sc + "/bbox_pred/weights"]) + w[sc + "/bbox_pred/biases"]
    stds = np.tile(np.asarray(o["bbox_stds"], dtype=np.float64), num_classes)
    means = np.tile(np.asarray(o["bbox_means"], dtype=np.float64), num_classes)
    bbox = (bbox * stds.astype(F) + means.astype(F)).astype(F)    # network.py:431-432 on a fp32 tensor
    return cls_score.astype(F), cls_prob.astype(F), bbox


def test_image(net, w, blob, im_info, num_classes, o=None, tap=None):
    """Network.test_image: -> dict(cls_score, cls_prob, bbox_pred, rois, + every stage tensor)."""
    o = o or opts()
    im_info = np.asarray(im_info, dtype=F)
    st = {}
    st["feat"] = N.image_to_head(net, w, blob, tap)
    st["rpn"], st["rpn_cls_score"], st["rpn_bbox_pred"] = rpn_head(net, w, st["feat"])
    st["rpn_scores"], st["rpn_props"], st["anchors"] = rpn_decode(st["rpn_cls_score"], st["rpn_bbox_pred"], im_info, o)
    st["rois"], st["roi_scores"], st["roi_keep"] = proposals(st["rpn_scores"], st["rpn_props"], o)
    st["pool5"] = crop_pool(net, st["feat"], st["rois"], o)
    st["fc7"] = N.head_to_tail(net, w, st["pool5"], tap)
    st["cls_score"], st["cls_prob"], st["bbox_pred"] = region_classification(net, w, st["fc7"], num_classes, o)
    return st


def get_image_blob(im_bgr_u8, o=None):
    """test.py:_get_image_blob for one scale: -> blob [1,H,W,3] fp32, im_scale (python float)."""
    import cv2
    o = o or opts()
    im = im_bgr_u8.astype(F, copy=True)
    im -= np.asarray(o["pixel_means"], dtype=np.float64).reshape(1, 1, 3)
    short, long_ = min(im.shape[:2]), max(im.shape[:2])
    scale = float(o["test_scale"]) / float(short)
    if np.round(scale * long_) > o["test_max_size"]:
        scale = float(o["test_max_size"]) / float(long_)
    im = cv2.resize(im, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
    return np.ascontiguousarray(im[None], dtype=F), scale


def im_detect_post(rois, cls_prob, bbox_pred, im_scale, orig_h, orig_w):
    """test.py:95-107 with the division pinned to fp32 (era numpy)."""
    boxes = (rois[:, 1:5] / F(im_scale)).astype(F)
    pred = B.clip_one_sided(B.decode(boxes, bbox_pred), orig_h, orig_w)
    return cls_prob.reshape(cls_prob.shape[0], -1), pred


def test_net_post(scores, boxes, o=None, thresh=0.0):
    """test.py:162-180 -> list over classes (index 0 = background = empty) of fp32 [k,5]."""
    o = o or opts()
    C = scores.shape[1]
    out = [np.zeros((0, 5), F)]
    for j in range(1, C):
        inds = np.where(scores[:, j] > thresh)[0]
        dets = np.hstack([boxes[inds, 4 * j:4 * j + 4], scores[inds, j][:, None]]).astype(F)
        keep = NMS.nms_plus1_c(dets, o["nms_thresh"], inclusive=not o["use_gpu_nms"])
        out.append(dets[keep])
    mpi = o["max_per_image"]
    if mpi > 0:
        allsc = np.hstack([d[:, 4] for d in out[1:]]) if C > 1 else np.zeros(0, F)
        if allsc.shape[0] > mpi:
            th = np.sort(allsc)[-mpi]
            out = [out[0]] + [d[d[:, 4] >= th] for d in out[1:]]
    return out
