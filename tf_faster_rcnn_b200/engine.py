"""Host-side executor of the TEST-mode graph on one GPU.

A `Tape` records, for one input shape, the ordered list of kernel launches (closures over
pre-built conv plans / static device buffers).  `Tape.run()` enqueues them on the current CUDA
stream; `ShapePlan` wraps a tape in a CUDA graph so that one image = one graph launch
(no tracing compiler: the graph is a recording of the explicit launches below).

Layer semantics follow lib/nets/network.py:233-262 (graph order), :323-378 (heads); the three
backbones emit their layers through the Tape from lib/nets/{vgg16,resnet_v1,mobilenet_v1}.py's
counterparts in tf_faster_rcnn_b200/lib/nets/.
"""
import numpy as np
import torch

from . import _native as N
from . import ops

F = np.float32


def bn_fold(gamma, beta, mean, var, eps):
    """tf.nn.batch_normalization (inference): y = x*inv + (beta - mean*inv), inv = gamma*rsqrt(var+eps)."""
    inv = (gamma.astype(F) * (F(1.0) / np.sqrt(var.astype(F) + F(eps)))).astype(F)
    return inv, (beta.astype(F) - mean.astype(F) * inv).astype(F)


class Weights:
    """TF-variable-name -> numpy store plus a cache of device-packed layers (packed once per network)."""

    def __init__(self, tensors):
        self.t = tensors
        self._packed = {}
        self._dev = {}

    def __getitem__(self, k):
        return self.t[k]

    def __contains__(self, k):
        return k in self.t

    def dev(self, key, arr_fn):
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(np.ascontiguousarray(arr_fn(), dtype=F)).cuda()
        return self._dev[key]

    def scale_shift(self, name, bn_eps):
        """epilogue vectors of layer `name`: BatchNorm fold when bn_eps is given, else (None, biases|None)."""
        if bn_eps is not None:
            p = name + "/BatchNorm/"
            return bn_fold(self.t[p + "gamma"], self.t[p + "beta"], self.t[p + "moving_mean"], self.t[p + "moving_variance"], bn_eps)
        b = self.t.get(name + "/biases")
        return None, (None if b is None else b.astype(F))

    def packed_conv(self, name, bn_eps=None, w_key="/weights"):
        if name not in self._packed:
            sc, sh = self.scale_shift(name, bn_eps)
            self._packed[name] = ops.PackedConv(self.t[name + w_key], sc, sh)
        return self._packed[name]

    def packed_custom(self, key, build):
        """build() -> (w_hwio, scale, shift) for fused layers (RPN heads, cls+bbox)."""
        if key not in self._packed:
            w, sc, sh = build()
            self._packed[key] = ops.PackedConv(w, sc, sh)
        return self._packed[key]


class Tape:
    def __init__(self, weights):
        self.w = weights
        self.steps = []        # (label, callable)
        self.flops = 0.0       # algorithmic 2*MAC of every conv/FC layer recorded
        self.conv_flops = 0.0  # ... of the tcgen05 implicit-GEMM launches only
        self.bufs = []
        self.conv_plans = []   # ops.ConvPlan objects, in launch order

    def new(self, *shape, dtype=torch.float32):
        t = torch.empty(shape, dtype=dtype, device="cuda")
        self.bufs.append(t)
        return t

    def add(self, label, fn):
        self.steps.append((label, fn))

    def run(self):
        for _, fn in self.steps:
            fn()

    # ---- dense layers -------------------------------------------------------------------------------
    def conv(self, x, name, stride=1, mode="SAME", act=N.ACT_RELU, bn_eps=None, residual=None, packed=None):
        pc = packed if packed is not None else self.w.packed_conv(name, bn_eps)
        n, h, w, _ = x.shape
        ho, wo, pt, pl = ops.conv_out_hw(h, w, pc.kh, stride, mode)
        out = self.new(n, ho, wo, pc.cout)
        plan = ops.ConvPlan(x, pc, out, stride, pt, pl, act, residual)
        self.conv_plans.append(plan)
        self.add("conv:" + name, plan.run)
        fl = 2.0 * n * ho * wo * pc.cout * pc.kh * pc.kw * pc.cin
        self.flops += fl
        self.conv_flops += fl
        return out

    def fc(self, x2d, name, act=N.ACT_NONE, packed=None):
        r, k = x2d.shape
        out = self.conv(x2d.view(1, 1, r, k), name, 1, "SAME", act, None, None, packed)
        return out.view(r, -1)

    def conv_first(self, x, name, k, stride, mode, act, bn_eps=None):
        sc, sh = self.w.scale_shift(name, bn_eps)
        wd = self.w.dev(name + "/weights", lambda: self.w[name + "/weights"])
        scd = None if sc is None else self.w.dev(name + "/scale", lambda: sc)
        shd = None if sh is None else self.w.dev(name + "/shift", lambda: sh)
        n, h, w, _ = x.shape
        ho, wo, pt, pl = ops.conv_out_hw(h, w, k, stride, mode)
        cout = int(self.w[name + "/weights"].shape[3])
        out = self.new(n, ho, wo, cout)
        self.add("conv_first:" + name, lambda: ops.conv_first(x, wd, scd, shd, out, k, stride, pt, pl, act))
        self.flops += 2.0 * n * ho * wo * cout * k * k * 3
        return out

    def depthwise(self, x, name, stride, act, bn_eps):
        sc, sh = self.w.scale_shift(name, bn_eps)
        c = x.shape[3]
        wd = self.w.dev(name + "/dw", lambda: self.w[name + "/depthwise_weights"].reshape(3, 3, c))
        scd = self.w.dev(name + "/scale", lambda: sc)
        shd = self.w.dev(name + "/shift", lambda: sh)
        n, h, w, _ = x.shape
        ho, wo, pt, pl = ops.conv_out_hw(h, w, 3, stride, "SAME" if stride == 1 else "EXPLICIT")
        out = self.new(n, ho, wo, c)
        self.add("depthwise:" + name, lambda: ops.depthwise3x3(x, wd, scd, shd, out, stride, pt, pl, act))
        return out

    def max_pool(self, x, k, stride, mode):
        """mode 'SAME' (TF: padded cells never win) | 'ZEROPAD1' (tf.pad 1 + VALID) | 'VALID'."""
        n, h, w, c = x.shape
        if mode == "SAME":
            ho, wo = -(-h // stride), -(-w // stride)
            pt, _ = ops.same_pads(h, k, stride); pl, _ = ops.same_pads(w, k, stride)
            neg = True
        elif mode == "ZEROPAD1":
            ho, wo, pt, pl, neg = (h + 2 - k) // stride + 1, (w + 2 - k) // stride + 1, 1, 1, False
        else:
            ho, wo, pt, pl, neg = (h - k) // stride + 1, (w - k) // stride + 1, 0, 0, True
        out = self.new(n, ho, wo, c)
        self.add("max_pool", lambda: ops.max_pool(x, out, k, stride, pt, pl, neg))
        return out

    def spatial_mean(self, x):
        out = self.new(x.shape[0], x.shape[3])
        self.add("spatial_mean", lambda: ops.spatial_mean(x, out))
        return out


class ShapePlan:
    """Everything needed to run one blob shape: static input/output buffers, the tape, its CUDA graph."""

    def __init__(self, net, h, w, use_graph=True):
        self.net, self.h, self.w = net, h, w
        cfgd = net.options
        wts = net.weights
        t = Tape(wts)
        self.tape = t
        self.image = t.new(1, h, w, 3)
        self.im_info = np.zeros(3, F)
        A = net.num_anchors
        C = net.num_classes
        sc = net.scope
        # ---- backbone -----------------------------------------------------------------------------------
        feat = net._image_to_head(t, self.image)
        self.feat = feat
        _, fh, fw, cb = feat.shape
        assert fh == -(-h // 16) and fw == -(-w // 16), "feature map %dx%d does not match ceil(H/16) x ceil(W/16)" % (fh, fw)
        # ---- RPN (network.py:323-359): 3x3 conv + ONE fused 1x1 for cls(2A) | pad | bbox(4A) ---------------
        rpn = t.conv(feat, sc + "/rpn_conv/3x3", 1, "SAME", N.ACT_RELU)
        dcol = (2 * A + 3) // 4 * 4
        ld = (dcol + 4 * A + 3) // 4 * 4

        def fused_rpn():
            cin = int(wts[sc + "/rpn_cls_score/weights"].shape[2])
            wf = np.zeros((1, 1, cin, ld), F); bf = np.zeros(ld, F)
            wf[..., :2 * A] = wts[sc + "/rpn_cls_score/weights"]; bf[:2 * A] = wts[sc + "/rpn_cls_score/biases"]
            wf[..., dcol:dcol + 4 * A] = wts[sc + "/rpn_bbox_pred/weights"]; bf[dcol:dcol + 4 * A] = wts[sc + "/rpn_bbox_pred/biases"]
            return wf, None, bf
        rpn_out = t.conv(rpn, sc + "/rpn_heads", 1, "SAME", N.ACT_NONE, packed=wts.packed_custom(sc + "/rpn_heads", fused_rpn))
        self.rpn_out, self.rpn_dcol = rpn_out, dcol
        nanch = fh * fw * A
        self.rpn_scores = t.new(nanch); self.rpn_props = t.new(nanch, 4)
        base = wts.dev("base_anchors/%s" % (net.anchor_key,), lambda: net.base_anchors)
        im_hw = (float(h), float(w))
        t.add("rpn_decode", lambda: ops.rpn_decode(rpn_out.view(fh * fw, ld), dcol, base, A, fh, fw, im_hw[0], im_hw[1],
                                                   self.rpn_scores, self.rpn_props))
        self.order = t.new(nanch, dtype=torch.int32); self.sorted_scores = t.new(nanch)
        sort_ws = ops.sort_workspace(nanch); t.bufs.append(sort_ws)
        t.add("sort_desc", lambda: ops.sort_desc(self.rpn_scores, self.order, self.sorted_scores, sort_ws))
        # ---- proposals (proposal_layer_tf | proposal_layer | proposal_top_layer) ---------------------------
        if cfgd["test_mode"] == "top":
            R, pre, thr, flags = cfgd["rpn_top_n"], 0, -1.0, 0
        elif cfgd["use_e2e_tf"]:
            R, pre, thr, flags = cfgd["rpn_post_nms_top_n"], 0, float(F(cfgd["rpn_nms_thresh"])), N.NMS_MODE_TF
        else:
            R, pre = cfgd["rpn_post_nms_top_n"], cfgd["rpn_pre_nms_top_n"]
            thr, flags = nms_threshold(cfgd["rpn_nms_thresh"], cfgd["use_gpu_nms"])
        self.R = R
        self.rois = t.new(R, 5); self.roi_scores = t.new(R)
        self.roi_keep = t.new(R, dtype=torch.int32); self.num_rois = t.new(1, dtype=torch.int32)
        t.add("proposals", lambda: ops.proposals(self.rpn_props, self.rpn_scores, self.order, pre, R, thr, flags, self.rois,
                                                 self.roi_scores, self.roi_keep, self.num_rois))
        # ---- RoI pooling (network.py:141-157 / resnet_v1.py:55-76) ---------------------------------------
        P = cfgd["pooling_size"]
        pre_pool = net.crop_pre_pool()
        self.pool5 = t.new(R, P, P, cb)
        t.add("crop_pool", lambda: ops.crop_pool(feat, self.rois, P, pre_pool, self.pool5))
        # ---- per-RoI head + fused cls_score|bbox_pred FC (network.py:361-378) ------------------------------
        fc7 = net._head_to_tail(t, self.pool5)
        self.fc7 = fc7

        ld_head = (5 * C + 3) // 4 * 4            # zero-padded to a multiple of 4 columns: vector stores + split-K apply

        def fused_cls():
            wc, wb = wts[sc + "/cls_score/weights"], wts[sc + "/bbox_pred/weights"]
            wf = np.zeros((1, 1, wc.shape[0], ld_head), F); bf = np.zeros(ld_head, F)
            wf[0, 0, :, :C] = wc; wf[0, 0, :, C:5 * C] = wb
            bf[:C] = wts[sc + "/cls_score/biases"]; bf[C:5 * C] = wts[sc + "/bbox_pred/biases"]
            return wf, None, bf
        self.head_out = t.fc(fc7, sc + "/cls_bbox", N.ACT_NONE, packed=wts.packed_custom(sc + "/cls_bbox", fused_cls))
        self.cls_score = t.new(R, C); self.cls_prob = t.new(R, C); self.bbox_pred = t.new(R, 4 * C)
        stds, means = cfgd["bbox_stds"], cfgd["bbox_means"]
        t.add("cls_finish", lambda: ops.cls_finish(self.head_out, C, stds, means, self.cls_score, self.cls_prob, self.bbox_pred))
        self.n_test_image_steps = len(t.steps)
        # ---- im_detect / test_net tail on device (test.py:95-107,162-180) ---------------------------------
        self.pred_boxes = t.new(R, 4 * C)
        self.post = dict(im_scale=1.0, orig_h=h, orig_w=w)
        self.max_det = 2 * cfgd["max_per_image"] + 56 if cfgd["max_per_image"] > 0 else R * (C - 1)
        self.det = t.new(self.max_det, 6); self.ndet = t.new(1, dtype=torch.int32)
        self.keep = t.new(C, R, dtype=torch.int32); self.keep_cnt = t.new(C, dtype=torch.int32)
        self.keep_score = t.new(C, R)
        self.graphs = {}
        self.use_graph = use_graph

    # the post-processing steps depend on per-image scalars (scale, original size): they are enqueued
    # directly after the graph instead of being baked into it.
    def _post(self, im_scale, orig_h, orig_w, detect):
        net = self.net
        C = net.num_classes
        ops.bbox_decode(self.rois, self.bbox_pred, C, im_scale, orig_h, orig_w, self.pred_boxes)
        if detect:
            o = net.options
            thr, flags = nms_threshold(o["nms_thresh"], o["use_gpu_nms"])
            ops.detect_post(self.cls_prob, self.pred_boxes, self.num_rois, C, o["score_thresh"], thr, flags, o["max_per_image"],
                            self.det, self.ndet, self.keep, self.keep_cnt, self.keep_score)

    def launch(self, im_scale=1.0, orig_h=None, orig_w=None, post=False, detect=False):
        """Enqueue one image (input already in self.image) on the current stream."""
        if self.use_graph:
            g = self.graphs.get("main")
            if g is None:
                self.tape.run()                       # warm-up: function attributes, lazy allocations
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.tape.run()
                self.graphs["main"] = g
            g.replay()
        else:
            self.tape.run()
        if post or detect:
            self._post(float(F(im_scale)), int(orig_h if orig_h is not None else self.h), int(orig_w if orig_w is not None else self.w), detect)


def nms_threshold(thresh, use_gpu_nms):
    """(fp32 threshold, flags) reproducing the reference's two '+1' predicates:
    cpu_nms compares the fp32 overlap with a DOUBLE threshold using >= (cpu_nms.pyx:17,65)  <=> ovr >= ceil32(t);
    gpu_nms compares with float(t) using > (nms_kernel.cu:34,71)."""
    t32 = F(thresh)
    if use_gpu_nms:
        return float(t32), N.NMS_MODE_GPU_NMS
    if float(t32) < float(thresh):
        t32 = np.nextafter(t32, F(np.inf))
    return float(t32), N.NMS_MODE_CPU_NMS
