"""One small invocation of the hot path on cuda:0, checked against the oracle (called by
__graft_entry__.smoke() only; the oracle import lives here because smoke is a checker)."""


def run(np, torch):
    from oracle import pipeline as P, nms as ONMS
    from . import ops, synth, paths, _native as N
    paths.add_lib_path()
    from model.config import cfg
    from nets.resnet_v1 import resnetv1
    F = np.float32
    # flagship model (ResNet-101, COCO anchors, 81 classes) on a small blob: every stage of the path runs once
    C, scales, hw = 81, (4, 8, 16, 32), (160, 224)
    cfg.TEST.HAS_RPN = True
    net = resnetv1(num_layers=101)
    net.create_architecture("TEST", C, tag="default", anchor_scales=scales, anchor_ratios=(0.5, 1, 2))
    w = synth.make("res101", C, 3 * len(scales))
    net.load_weights(w)
    blob = synth.synthetic_blob(*hw)
    im_info = np.array([hw[0], hw[1], 1.0], F)
    cls_score, cls_prob, bbox_pred, rois = net.test_image(None, blob, im_info)
    st = P.test_image("res101", w, blob, im_info, C, P.opts(anchor_scales=scales))
    plan = net.plan_for(*hw)
    keep = plan.roi_keep.cpu().numpy()[:rois.shape[0]]
    common, ia, ib = np.intersect1d(keep, st["roi_keep"], return_indices=True)
    assert len(common) >= 0.97 * len(st["roi_keep"]), "RoI sets differ: %d common of %d" % (len(common), len(st["roi_keep"]))
    e_prob = float(np.abs(cls_prob[ia] - st["cls_prob"][ib]).max())
    e_box = float(np.abs(bbox_pred[ia] - st["bbox_pred"][ib]).max())
    assert e_prob < 1e-4 and e_box < 1e-4, (e_prob, e_box)
    det, _ = net.detect(blob, im_info, hw)
    assert det.shape[1] == 6 and 0 < det.shape[0] <= plan.max_det
    # NMS through the `_nms`-compatible entry
    rng = np.random.default_rng(0)
    n = 500
    xy = rng.uniform(0, 400, (n, 2)); wh = rng.uniform(10, 120, (n, 2))
    dets = np.hstack([xy, xy + wh, rng.random((n, 1))]).astype(F)
    order = ONMS.argsort_desc(dets[:, 4])
    got = order[ops.nms_host(dets[order], float(ONMS.thresh_f32(0.3, True)), N.NMS_MODE_CPU_NMS)]
    assert np.array_equal(got, ONMS.nms_plus1_c(dets, 0.3, True)), "nms parity"
    print("smoke: ResNet-101 %dx%d  rois %d (common %d)  cls_prob err %.2e  bbox_pred err %.2e  dets %d" %
          (hw[0], hw[1], rois.shape[0], len(common), e_prob, e_box, det.shape[0]))
    return e_prob
