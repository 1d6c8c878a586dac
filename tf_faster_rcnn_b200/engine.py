"""Host-side executor of the TEST-mode graph on one GPU.

A `Tape` records, for one input shape, the ordered list of kernel launches (closures over
pre-built conv plans / static device buffers).  `Tape.run()` enqueues them on the current CUDA
stream; `ShapePlan` wraps a tape in a CUDA graph so that one image = one graph launch
(no tracing compiler: the graph is a recording of the explicit launches below).

Layer semantics follow lib/nets/network.py:233-262 (graph order), :323-378 (heads); the three
backbones emit their layers through the Tape from lib/nets/{vgg16,resnet_v1,mobilenet_v1}.py's
counterparts in tf_faster_rcnn_b200/lib/nets/.
"""
import ctypes
import os

import numpy as np
import torch

from . import _native as N
from . import ops

F = np.float32
C_void = ctypes.c_void_p


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


class LaunchGraph:
    """A recorded launch sequence as an executable CUDA graph, captured and replayed through the C ABI (frcnn_graph_*): plain
    cudaStreamBeginCapture / cudaGraphLaunch on the stream the stages were enqueued on -- no framework graph object, no
    generator-state fill kernels per replay.  FRCNN_TORCH_GRAPH=1 selects torch.cuda.CUDAGraph instead (A/B, debugging)."""

    def __init__(self, fns):
        self._h = ctypes.c_void_p()
        self._torch = None
        if os.environ.get("FRCNN_TORCH_GRAPH") == "1":
            self._torch = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._torch):
                for fn in fns:
                    fn()
            return
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            N.check(N.lib().frcnn_graph_begin(C_void(side.cuda_stream)), "graph_begin")
            try:
                for fn in fns:
                    fn()
            finally:
                rc = N.lib().frcnn_graph_end(C_void(side.cuda_stream), ctypes.byref(self._h))
            N.check(rc, "graph_end")
        torch.cuda.current_stream().wait_stream(side)

    def replay(self):
        if self._torch is not None:
            self._torch.replay()
        else:
            N.check(N.lib().frcnn_graph_launch(self._h, ops._stream()), "graph_launch")

    def __del__(self):
        try:
            if self._h:
                N.lib().frcnn_graph_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


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


REC_HEADER = 8     # 4-byte words in front of every image's detection rows: word 0 = int32 detection count


class ShapePlan:
    """Everything needed to run `batch` images of one blob shape: static input/output buffers, the tape, its CUDA graphs.

    The reference graph is batch 1 (lib/nets/network.py:388); batch > 1 is the throughput mode: the
    backbone sees M = batch * H * W output pixels per layer (the 38x50 ResNet maps fill the 148 SMs without split-K), the
    per-RoI head sees batch * R RoIs, and the single-CTA proposal / NMS kernels run one CTA per image side by side.
    One image (or batch) = ONE graph replay: the im_detect / test_net tail is part of the graph, its per-image scalars
    (scale, original size) are read from `im_meta` on the device."""

    def __init__(self, net, h, w, batch=1, use_graph=True):
        self.net, self.h, self.w, self.batch = net, h, w, batch
        cfgd = net.options
        wts = net.weights
        t = Tape(wts)
        self.tape = t
        B = batch
        self.image = t.new(B, h, w, 3)
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
        self.nanch = nanch
        self.rpn_scores = t.new(B * nanch); self.rpn_props = t.new(B * nanch, 4)
        base = wts.dev("base_anchors/%s" % (net.anchor_key,), lambda: net.base_anchors)
        im_hw = (float(h), float(w))
        t.add("rpn_decode", lambda: ops.rpn_decode(rpn_out.view(B * fh * fw, ld), dcol, base, A, fh, fw, im_hw[0], im_hw[1],
                                                   self.rpn_scores, self.rpn_props, batch=B))
        self.order = t.new(B * nanch, dtype=torch.int32); self.sorted_scores = t.new(B * nanch)
        t.add("sort_desc", lambda: ops.sort_desc(self.rpn_scores, self.order, self.sorted_scores, None, batch=B))
        # ---- proposals (proposal_layer_tf | proposal_layer | proposal_top_layer) ---------------------------
        if cfgd["test_mode"] == "top":
            R, pre, thr, flags = cfgd["rpn_top_n"], 0, -1.0, 0
        elif cfgd["use_e2e_tf"]:
            R, pre, thr, flags = cfgd["rpn_post_nms_top_n"], 0, float(F(cfgd["rpn_nms_thresh"])), N.NMS_MODE_TF
        else:
            R, pre = cfgd["rpn_post_nms_top_n"], cfgd["rpn_pre_nms_top_n"]
            thr, flags = nms_threshold(cfgd["rpn_nms_thresh"], cfgd["use_gpu_nms"])
        self.R = R
        self.rois = t.new(B * R, 5); self.roi_scores = t.new(B * R)
        self.roi_keep = t.new(B * R, dtype=torch.int32); self.num_rois = t.new(B, dtype=torch.int32)
        t.add("proposals", lambda: ops.proposals(self.rpn_props, self.rpn_scores, self.order, pre, R, thr, flags, self.rois,
                                                 self.roi_scores, self.roi_keep, self.num_rois, batch=B))
        # ---- RoI pooling (network.py:141-157 / resnet_v1.py:55-76) ---------------------------------------
        P = cfgd["pooling_size"]
        pre_pool = net.crop_pre_pool()
        self.pool5 = t.new(B * R, P, P, cb)
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
        self.cls_score = t.new(B * R, C); self.cls_prob = t.new(B * R, C); self.bbox_pred = t.new(B * R, 4 * C)
        stds, means = cfgd["bbox_stds"], cfgd["bbox_means"]
        t.add("cls_finish", lambda: ops.cls_finish(self.head_out, C, stds, means, self.cls_score, self.cls_prob, self.bbox_pred))
        self.n_test_image_steps = len(t.steps)
        # ---- im_detect tail on the device (test.py:95-107): per-image (scale, orig_h, orig_w) live in im_meta --------------
        self.pred_boxes = t.new(B * R, 4 * C)
        self.im_meta = t.new(B, 3)
        # pinned staging for the meta rows: a ring, because the H2D copies are asynchronous and the host may already be
        # preparing the launch after next (submit_batch keeps two batches in flight)
        self.im_meta_ring = [torch.empty((B, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
        self.im_meta_turn = 0
        self.im_meta_ring[0][:] = torch.tensor([1.0, float(h), float(w)])
        self.im_meta.copy_(self.im_meta_ring[0])
        t.add("bbox_decode", lambda: ops.bbox_decode(self.rois, self.bbox_pred, C, self.im_meta, self.pred_boxes))
        self.n_im_detect_steps = len(t.steps)
        # ---- test_net tail (test.py:162-180): built on first use for the options in force (_ensure_post) ---------------------
        self.keep = t.new(B, C, R, dtype=torch.int32); self.keep_cnt = t.new(B, C, dtype=torch.int32)
        self.keep_score = t.new(B, C, R)
        self.post_ws = ops.detect_post_workspace(R, C, B); t.bufs.append(self.post_ws)
        self.post_key = None
        self.recs = [None, None]       # two record buffers; `double_buffer` makes consecutive detect launches alternate
        self.rec = self.det = self.ndet = None   # ... views of the buffer the LAST detect launch wrote
        self.post_steps = [None, None]
        self.double_buffer = False
        self.slot = 0
        self.max_det = 0
        self.graphs = {}
        self.use_graph = use_graph

    def nbytes(self):
        return sum(b.numel() * b.element_size() for b in self.tape.bufs)

    def release(self):
        """Drop graphs, conv plans (TMA descriptors, split-K workspaces) and activation buffers (LRU eviction)."""
        self.graphs.clear()
        self.tape.steps = []
        self.tape.conv_plans = []
        self.tape.bufs = []
        self.post_steps = [None, None]
        self.recs = [None, None]
        self.rec = self.det = self.ndet = None

    def _ensure_post(self):
        """(Re)build the detection-record buffers and the post step for the current score / NMS thresholds and cap."""
        net = self.net
        o = net.options
        key = (float(o["score_thresh"]), float(o["nms_thresh"]), bool(o["use_gpu_nms"]), int(o["max_per_image"]))
        if key == self.post_key:
            return
        C, R, B = net.num_classes, self.R, self.batch
        mpi = key[3]
        # records: max_per_image survivors + head-room for ties at the threshold score; no cap -> every (roi, class) pair.
        # A record set that still does not fit is reported through ndet > max_det and raised on the host (never truncated).
        self.max_det = 2 * mpi + 56 if mpi > 0 else R * (C - 1)
        stride = REC_HEADER + self.max_det * 6
        thr, flags = nms_threshold(key[1], key[2])

        def make_post(rec):
            def post():
                N.check(N.lib().frcnn_detect_post(ops._p(self.cls_prob), ops._p(self.pred_boxes), ops._p(self.num_rois), R, B, C, key[0],
                                                  thr, flags, mpi, self.max_det, C_void(rec.data_ptr() + 4 * REC_HEADER), ops._p(rec), stride,
                                                  ops._p(self.keep), ops._p(self.keep_cnt), ops._p(self.keep_score), ops._p(self.post_ws),
                                                  self.post_ws.numel(), ops._stream()), "detect_post")
            return post
        self.recs = [ops.zeros((B, stride)) for _ in range(2)]
        self.post_steps = [make_post(r) for r in self.recs]
        self.post_key = key
        self._select(0)
        self.graphs.pop(("detect", 0), None); self.graphs.pop(("detect", 1), None)

    def _select(self, slot):
        self.slot = slot
        self.rec = self.recs[slot]
        self.det = [self.rec[b, REC_HEADER:].view(self.max_det, 6) for b in range(self.batch)]
        self.ndet = self.rec.view(torch.int32)[:, 0]

    def steps_for(self, mode):
        """mode: 'test_image' (network outputs), 'im_detect' (+ decoded boxes), 'detect' (+ per-class NMS, cap, records)."""
        if mode == "test_image":
            return [fn for _, fn in self.tape.steps[:self.n_test_image_steps]]
        fns = [fn for _, fn in self.tape.steps[:self.n_im_detect_steps]]
        if mode == "detect":
            self._ensure_post()
            fns.append(self.post_steps[self.slot])
        return fns

    def set_meta(self, rows):
        """rows: per image (im_scale, orig_h, orig_w); staged in pinned memory, copied on the launching stream."""
        self.im_meta_turn = (self.im_meta_turn + 1) & 3
        host = self.im_meta_ring[self.im_meta_turn]
        for b, (s, oh, ow) in enumerate(rows):
            host[b, 0] = float(F(s)); host[b, 1] = float(oh); host[b, 2] = float(ow)
        self.im_meta.copy_(host, non_blocking=True)

    def launch(self, im_scale=1.0, orig_h=None, orig_w=None, post=False, detect=False, meta=None):
        """Enqueue one batch (inputs already in self.image) on the current stream.  meta: per-image (scale, orig_h, orig_w)
        rows; the scalar arguments describe every image of the batch when meta is None."""
        mode = "detect" if detect else ("im_detect" if post else "test_image")
        if mode != "test_image":
            if meta is None:
                meta = [(im_scale, orig_h if orig_h is not None else self.h, orig_w if orig_w is not None else self.w)] * self.batch
            self.set_meta(meta)
        gkey = mode
        if mode == "detect":
            self._ensure_post()
            self._select(self.slot ^ 1 if self.double_buffer else 0)
            gkey = ("detect", self.slot)
        if self.use_graph:
            g = self.graphs.get(gkey)
            if g is None:
                fns = self.steps_for(mode)
                for fn in fns:                        # warm-up: function attributes, lazy allocations
                    fn()
                torch.cuda.current_stream().synchronize()
                g = LaunchGraph(fns)
                self.graphs[gkey] = g
            g.replay()
        else:
            for fn in self.steps_for(mode):
                fn()

    def records(self):
        """Host copy of the detection records of the batch after a 'detect' launch: list of [n,6] arrays (one D2H copy)."""
        return split_host_records(self.rec.cpu(), self.max_det)


def split_host_records(host, max_det):
    """host: CPU float32 tensor [B, REC_HEADER + max_det*6] -> list of [n,6] numpy arrays; raises when a record set did not fit."""
    counts = host.view(torch.int32)[:, 0].numpy()
    out = []
    for b in range(host.shape[0]):
        n = int(counts[b])
        if n > max_det:
            raise RuntimeError("image %d of the batch produced %d detections but the record buffer holds %d "
                               "(score ties beyond the max_per_image head-room)" % (b, n, max_det))
        out.append(host[b, REC_HEADER:REC_HEADER + n * 6].view(n, 6).numpy().copy())
    return out


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
