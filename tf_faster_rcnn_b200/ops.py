"""Stage-level Python entry points over the C ABI.  torch is used ONLY for device memory
(`torch.empty(..., device='cuda')`, `.data_ptr()`) and the current CUDA stream; every op below
launches hand-written sm_100a kernels from libfrcnn_b200.so.  No CPU fallbacks.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t):
    assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous(), "expected contiguous cuda fp32"
    return t


def zeros(shape, dtype=torch.float32):
    """Device buffer cleared with cudaMemsetAsync on the current stream (no framework fill kernel on the path)."""
    t = torch.empty(shape, dtype=dtype, device="cuda")
    N.check(N.lib().frcnn_zero_async(_p(t), t.numel() * t.element_size(), _stream()), "zero_async")
    return t


def same_pads(n, k, s):
    """TF 'SAME' padding (before, after) for one dimension."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv_impl():
    """Dense-kernel arithmetic (FRCNN_CONV_IMPL): 'f16' = FP16x3, fp32-grade (default); 'tf32' = the r01 TF32x3 kernel (A/B
    measurements); 'f16x1' = THROUGHPUT mode, plain fp16 operands with fp32 accumulation (NOT fp32-grade: ~3e-4 per layer)."""
    return {"tf32": N.CONV_TF32X3, "f16x1": N.CONV_F16X1}.get(os.environ.get("FRCNN_CONV_IMPL", "f16"), N.CONV_F16X3)


def weight_exponent(w):
    """wexp with max|w| * 2^wexp in [2^13, 2^14): the fp16 hi plane then spans 27 binades below the layer's largest
    weight before going subnormal (and even those keep 2^-35 * max|w| through the lo plane)."""
    m = float(np.max(np.abs(w))) if w.size else 0.0
    if not np.isfinite(m) or m <= 0.0:
        return 0
    return int(13 - np.floor(np.log2(m)))


class PackedConv:
    """Device-resident K-major hi/lo weight planes + epilogue vectors of one conv / FC layer."""

    def __init__(self, w_hwio, scale=None, shift=None, impl=None):
        wnp = np.ascontiguousarray(w_hwio, dtype=np.float32)
        w = torch.from_numpy(wnp).cuda()
        if w.dim() == 2:                     # FC [in,out] == 1x1 conv
            w = w.view(1, 1, w.shape[0], w.shape[1])
        self.kh, self.kw, self.cin, self.cout = (int(v) for v in w.shape)
        ktot = self.kh * self.kw * self.cin
        self.impl = conv_impl() if impl is None else impl
        if self.impl in (N.CONV_F16X3, N.CONV_F16X1):
            self.wexp = weight_exponent(wnp)
            self.out_mult = float(np.ldexp(1.0, -self.wexp))
            self.w_hi = torch.empty((self.cout, ktot), dtype=torch.float16, device="cuda")
            self.w_lo = torch.empty_like(self.w_hi)
            N.check(N.lib().frcnn_pack_conv_weights(_p(w), _p(self.w_hi), _p(self.w_lo), self.kh, self.kw, self.cin,
                                                    self.cout, self.wexp, _stream()), "pack_conv_weights")
        else:
            self.wexp, self.out_mult = 0, 1.0
            self.w_hi = torch.empty((self.cout, ktot), dtype=torch.float32, device="cuda")
            self.w_lo = torch.empty_like(self.w_hi)
            N.check(N.lib().frcnn_pack_conv_weights_tf32(_p(w), _p(self.w_hi), _p(self.w_lo), self.kh, self.kw, self.cin,
                                                         self.cout, _stream()), "pack_conv_weights_tf32")
        torch.cuda.current_stream().synchronize()
        self.scale = None if scale is None else torch.as_tensor(np.ascontiguousarray(scale), dtype=torch.float32).cuda()
        self.shift = None if shift is None else torch.as_tensor(np.ascontiguousarray(shift), dtype=torch.float32).cuda()


class ConvPlan:
    """frcnn_conv_plan bound to fixed input/output/residual buffers (TMA descriptors hold raw pointers)."""

    def __init__(self, x, pc, out, stride=1, pad_t=0, pad_l=0, act=N.ACT_NONE, residual=None, block_n=0, kb_per_chunk=0, split_k=0):
        _f32(x); _f32(out)
        n, h, w, cin = x.shape
        assert cin == pc.cin, (cin, pc.cin)
        no, ho, wo, co = out.shape
        assert no == n and co == pc.cout
        d = N.ConvDesc(_p(x), _p(pc.w_hi), _p(pc.w_lo), _p(pc.scale), _p(pc.shift), _p(residual), _p(out),
                       n, h, w, cin, pc.cout, pc.kh, pc.kw, stride, pad_t, pad_l, ho, wo, act, block_n, kb_per_chunk, split_k,
                       pc.impl, pc.out_mult)
        self._h = C.c_void_p()
        N.check(N.lib().frcnn_conv_plan_create(C.byref(self._h), C.byref(d)), "conv_plan_create")
        self._keep = (x, pc, out, residual)
        self.flops = 2.0 * n * ho * wo * pc.cout * pc.kh * pc.kw * pc.cin

    def run(self):
        N.check(N.lib().frcnn_conv_plan_run(self._h, _stream()), "conv_plan_run")

    def info(self):
        v = [C.c_int() for _ in range(8)]
        N.check(N.lib().frcnn_conv_plan_info(self._h, *[C.byref(a) for a in v]), "conv_plan_info")
        return dict(zip(["block_n", "tile_n", "tile_h", "tile_w", "grid_m", "grid_n", "splits", "smem"], [a.value for a in v]))

    def __del__(self):
        try:
            if self._h:
                N.lib().frcnn_conv_plan_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def conv_out_hw(h, w, k, stride, mode):
    """mode 'SAME' (TF) or 'EXPLICIT' (slim conv2d_same: pad (k-1)//2 before, rest after, then VALID)."""
    if mode == "SAME":
        pt, _ = same_pads(h, k, stride)
        pl, _ = same_pads(w, k, stride)
        return -(-h // stride), -(-w // stride), pt, pl
    pb = (k - 1) // 2
    return (h + k - 1 - k) // stride + 1, (w + k - 1 - k) // stride + 1, pb, pb


def conv_first(x, w_hwio_dev, scale, shift, out, k, stride, pad_t, pad_l, act):
    n, h, w, _ = x.shape
    _, ho, wo, co = out.shape
    N.check(N.lib().frcnn_conv_first(_p(_f32(x)), _p(w_hwio_dev), _p(scale), _p(shift), _p(_f32(out)), n, h, w, co, k, stride,
                                     pad_t, pad_l, ho, wo, act, _stream()), "conv_first")


def depthwise3x3(x, w_dev, scale, shift, out, stride, pad_t, pad_l, act):
    n, h, w, c = x.shape
    _, ho, wo, _ = out.shape
    N.check(N.lib().frcnn_depthwise3x3(_p(_f32(x)), _p(w_dev), _p(scale), _p(shift), _p(_f32(out)), n, h, w, c, stride, pad_t,
                                       pad_l, ho, wo, act, _stream()), "depthwise3x3")


def max_pool(x, out, k, stride, pad_t, pad_l, pad_is_neg_inf):
    n, h, w, c = x.shape
    _, ho, wo, _ = out.shape
    N.check(N.lib().frcnn_max_pool(_p(_f32(x)), _p(_f32(out)), n, h, w, c, k, stride, pad_t, pad_l, ho, wo,
                                   int(pad_is_neg_inf), _stream()), "max_pool")


def spatial_mean(x, out):
    r, hw, c = x.shape[0], x.shape[1] * x.shape[2], x.shape[3]
    N.check(N.lib().frcnn_spatial_mean(_p(_f32(x)), _p(_f32(out)), r, hw, c, _stream()), "spatial_mean")


def preprocess(img_u8_dev, means3, fx, fy, blob):
    """uint8 BGR [h0,w0,3] (device) -> mean-subtracted, bilinearly resized fp32 blob [1,H,W,3] (device)."""
    h0, w0, _ = img_u8_dev.shape
    _, H, W, _ = blob.shape
    m = (C.c_double * 3)(*[float(v) for v in means3])
    N.check(N.lib().frcnn_preprocess(C.c_void_p(img_u8_dev.data_ptr()), h0, w0, m, float(fx), float(fy), _p(blob), H, W, _stream()),
            "preprocess")


def rpn_decode(rpn_out, delta_col, base_anchors, num_anchors, fh, fw, im_h, im_w, scores, props, feat_stride=16, batch=1):
    """rpn_out: [batch*fh*fw, ld] rows of the fused RPN head."""
    ld = rpn_out.shape[-1]
    N.check(N.lib().frcnn_rpn_decode(_p(_f32(rpn_out)), ld, delta_col, _p(base_anchors), num_anchors, batch, fh, fw, feat_stride,
                                     float(im_h), float(im_w), _p(scores), _p(props), _stream()), "rpn_decode")


def sort_workspace(n):
    return torch.empty(int(N.lib().frcnn_sort_workspace_bytes(n)), dtype=torch.uint8, device="cuda")


def sort_desc(keys, order, sorted_keys, workspace=None, batch=1):
    """`batch` segments of keys.numel() // batch keys each; order = segment-local indices."""
    n = keys.numel() // batch
    N.check(N.lib().frcnn_sort_desc(_p(_f32(keys)), n, batch, _p(order), _p(sorted_keys), _p(workspace),
                                    0 if workspace is None else workspace.numel(), _stream()), "sort_desc")


def proposals(props, scores, order, pre_nms_top_n, post_nms_top_n, thresh, flags, rois, roi_scores, keep, num, batch=1):
    n = scores.numel() // batch
    N.check(N.lib().frcnn_proposals(_p(props), _p(scores), _p(order), n, batch, pre_nms_top_n, post_nms_top_n, float(thresh), flags,
                                    _p(rois), _p(roi_scores), _p(keep), _p(num), _stream()), "proposals")


def crop_pool(feat, rois, pooled, pre_pool, out):
    """rois[:, 0] = index of the image (of feat's batch dimension) the box is cut from."""
    b, fh, fw, c = feat.shape
    r = rois.shape[0]
    N.check(N.lib().frcnn_crop_pool(_p(_f32(feat)), b, fh, fw, c, _p(_f32(rois)), r, pooled, int(pre_pool), _p(_f32(out)), _stream()),
            "crop_pool")


def cls_finish(head_out, num_classes, stds, means, cls_score, cls_prob, bbox_pred):
    r, ld = head_out.shape
    s4 = (C.c_float * 4)(*[float(v) for v in stds])
    m4 = (C.c_float * 4)(*[float(v) for v in means])
    N.check(N.lib().frcnn_cls_finish(_p(_f32(head_out)), ld, r, num_classes, s4, m4, _p(cls_score), _p(cls_prob), _p(bbox_pred),
                                     _stream()), "cls_finish")


def im_meta_tensor(rows):
    """[(im_scale, orig_h, orig_w), ...] -> device fp32 [batch, 3] (the per-image scalars bbox_decode reads)."""
    return torch.tensor([[float(np.float32(s)), float(h), float(w)] for s, h, w in rows], dtype=torch.float32).cuda()


def bbox_decode(rois, bbox_pred, num_classes, im_meta, pred_boxes):
    """im_meta: device fp32 [batch, 3] (im_scale, orig_h, orig_w); rois[:, 0] selects the row."""
    r = rois.shape[0]
    N.check(N.lib().frcnn_bbox_decode(_p(_f32(rois)), _p(_f32(bbox_pred)), r, num_classes, im_meta.shape[0], _p(_f32(im_meta)),
                                      _p(pred_boxes), _stream()), "bbox_decode")


def detect_post_workspace(r, num_classes, batch=1):
    return torch.empty(int(N.lib().frcnn_detect_post_workspace_bytes(r, num_classes, batch)), dtype=torch.uint8, device="cuda")


def detect_post(cls_prob, pred_boxes, num_rois, num_classes, score_thresh, nms_thresh, flags, max_per_image, det, ndet, keep,
                keep_cnt, keep_score, workspace=None, batch=1):
    """cls_prob [batch*r, C]; det [batch, max_det, 6] (or [max_det, 6] for batch 1); ndet int32 [batch] = TRUE counts
    (a count above max_det means the records did not fit)."""
    r = cls_prob.shape[0] // batch
    max_det = det.shape[-2]
    N.check(N.lib().frcnn_detect_post(_p(cls_prob), _p(pred_boxes), _p(num_rois), r, batch, num_classes, float(score_thresh),
                                      float(nms_thresh), flags, max_per_image, max_det, _p(det), _p(ndet), 0, _p(keep),
                                      _p(keep_cnt), _p(keep_score), _p(workspace), 0 if workspace is None else workspace.numel(),
                                      _stream()), "detect_post")


def nms_sorted_dev(boxes, thresh, flags, max_out, keep, num):
    N.check(N.lib().frcnn_nms_sorted_dev(_p(_f32(boxes)), boxes.shape[0], float(thresh), flags, max_out, _p(keep), _p(num),
                                         _stream()), "nms_sorted_dev")


def nms_host(sorted_dets, thresh, flags, device_id=-1):
    """`_nms`-compatible call on HOST arrays (already sorted by descending score).  device_id < 0: the current device."""
    d = np.ascontiguousarray(sorted_dets, dtype=np.float32)
    n = d.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int32)
    num = C.c_int(0)
    N.check(N.lib().frcnn_nms_host(keep.ctypes.data_as(N.ip), C.byref(num), d.ctypes.data_as(N.fp), n, d.shape[1] if n else 5,
                                   float(thresh), device_id, flags), "nms_host")
    return keep[:num.value].copy()
