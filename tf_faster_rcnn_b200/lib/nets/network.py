"""`Network` base class with the reference's call surface (lib/nets/network.py):
create_architecture :386-454, test_image :470-479, extract_head :464-467; subclasses implement
_image_to_head / _head_to_tail (:380-384).  Instead of building a TF graph, create_architecture records
the options; the first test_image for a blob shape builds a ShapePlan (device buffers, TMA-backed conv
plans, CUDA graph) and later calls replay it.  `sess` arguments are accepted and ignored.
"""
import os

import numpy as np
import torch

from model.config import cfg
from layer_utils.generate_anchors import generate_anchors
from tf_faster_rcnn_b200 import engine, _native

_REGISTRY = []   # networks created in this process (the tensorflow shim's Saver.restore walks it)


class Network(object):
    def __init__(self):
        self._feat_stride = [16, ]
        self._predictions = {}
        self._layers = {}
        self._scope = None
        self.weights = None
        self._plans = {}
        self.use_cuda_graph = True
        _REGISTRY.append(self)

    # ---- reference surface ---------------------------------------------------------------------------
    def create_architecture(self, mode, num_classes, tag=None, anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2)):
        assert tag is not None
        if mode != "TEST":
            raise NotImplementedError("only the TEST-mode (inference) graph exists in this build")
        self._mode, self._tag = mode, tag
        self._num_classes = num_classes
        self._anchor_scales, self._anchor_ratios = tuple(anchor_scales), tuple(anchor_ratios)
        self._num_scales, self._num_ratios = len(anchor_scales), len(anchor_ratios)
        self._num_anchors = self._num_scales * self._num_ratios
        self.base_anchors = generate_anchors(ratios=np.array(self._anchor_ratios), scales=np.array(self._anchor_scales)).astype(np.float32)
        self.anchor_key = "%s|%s" % (self._anchor_scales, self._anchor_ratios)
        if cfg.POOLING_MODE != "crop":
            raise NotImplementedError
        self.options = dict(
            test_mode=cfg.TEST.MODE, use_e2e_tf=bool(cfg.USE_E2E_TF), use_gpu_nms=bool(cfg.USE_GPU_NMS),
            rpn_nms_thresh=cfg.TEST.RPN_NMS_THRESH, rpn_pre_nms_top_n=cfg.TEST.RPN_PRE_NMS_TOP_N,
            rpn_post_nms_top_n=cfg.TEST.RPN_POST_NMS_TOP_N, rpn_top_n=cfg.TEST.RPN_TOP_N,
            pooling_size=cfg.POOLING_SIZE, resnet_max_pool=bool(cfg.RESNET.MAX_POOL),
            bbox_stds=tuple(cfg.TRAIN.BBOX_NORMALIZE_STDS), bbox_means=tuple(cfg.TRAIN.BBOX_NORMALIZE_MEANS),
            nms_thresh=cfg.TEST.NMS, max_per_image=100, score_thresh=0.0, rpn_channels=cfg.RPN_CHANNELS,
        )
        if self.options["test_mode"] not in ("nms", "top"):
            raise NotImplementedError
        self._plans = {}
        return {"rois": None}

    def test_image(self, sess, image, im_info):
        """-> (cls_score, cls_prob, bbox_pred, rois) host fp32 arrays, rois in blob-scale pixels."""
        plan = self._run(image, im_info)
        torch.cuda.current_stream().synchronize()
        r = int(plan.num_rois[0].item())
        out = (plan.cls_score[:r].cpu().numpy(), plan.cls_prob[:r].cpu().numpy(), plan.bbox_pred[:r].cpu().numpy(),
               plan.rois[:r].cpu().numpy())
        return out

    def extract_head(self, sess, image):
        plan = self._run(image, np.array([image.shape[1], image.shape[2], 1.0], np.float32))
        torch.cuda.current_stream().synchronize()
        return plan.feat.cpu().numpy()

    def _image_to_head(self, tape, image):
        raise NotImplementedError

    def _head_to_tail(self, tape, pool5):
        raise NotImplementedError

    # ---- device side -------------------------------------------------------------------------------------
    @property
    def num_anchors(self):
        return self._num_anchors

    @property
    def num_classes(self):
        return self._num_classes

    @property
    def scope(self):
        return self._scope

    def crop_pre_pool(self):
        """14x14 crop + 2x2 max pool (network.py:154-157) unless a subclass crops 7x7 directly."""
        return True

    def arch_name(self):
        """'vgg16' | 'res50' | 'res101' | 'res152' | 'mobile' (the --net names of tools/test_net.py:92-103)."""
        return {"vgg_16": "vgg16", "MobilenetV1": "mobile"}.get(self._scope) or "res%d" % self._num_layers

    def check_variables(self, tensors):
        """What tf.train.Saver.restore would reject: TEST-graph variables missing from `tensors` or of another shape
        (class count, anchor set, backbone).  Needs create_architecture() first.  -> list of messages."""
        from tf_faster_rcnn_b200 import synth
        return synth.check(self.arch_name(), tensors, self._num_classes, self._num_anchors,
                           rpn_channels=int(cfg.RPN_CHANNELS),
                           depth_multiplier=float(getattr(self, "_depth_multiplier", 1.0)))

    def load_weights(self, tensors, strict=False):
        """tensors: dict TF-variable-name -> numpy array (HWIO convs, [in,out] FCs, BatchNorm stats).
        strict: verify names and shapes against the architecture first (checkpoint restores) and raise ValueError."""
        if strict:
            problems = self.check_variables(tensors)
            if problems:
                raise ValueError("checkpoint does not match the %s TEST graph (%d classes, %d anchors):\n  %s"
                                 % (self.arch_name(), self._num_classes, self._num_anchors, "\n  ".join(problems)))
        self.weights = engine.Weights(dict(tensors))
        self._plans = {}

    MAX_PLANS = int(os.environ.get("FRCNN_MAX_PLANS", "6"))

    def plan_for(self, h, w, batch=1):
        """ShapePlan of blob shape (h, w) x batch.  A plan owns every layer's activation buffer, the TMA descriptors and the CUDA
        graphs of that shape (0.6-1.3 GB per image at 600x800..1000), so the cache is a small LRU: a dataset with hundreds of
        distinct shapes recycles plans instead of growing without bound (evicted buffers return to the caching allocator)."""
        if self.weights is None:
            raise RuntimeError("no weights loaded: call load_weights() / Saver.restore() before test_image")
        key = (int(h), int(w), int(batch))
        plan = self._plans.pop(key, None)
        if plan is None:
            dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
            _native.check(_native.lib().frcnn_check_device(dev), "check_device")   # fails loudly: no CPU fallback exists
            while len(self._plans) >= max(1, self.MAX_PLANS):
                old_key = next(iter(self._plans))
                old = self._plans.pop(old_key)
                torch.cuda.current_stream().synchronize()   # nothing of the evicted plan is still in flight
                old.release()
            plan = engine.ShapePlan(self, key[0], key[1], key[2], use_graph=self.use_cuda_graph)
        self._plans[key] = plan                             # most recently used last
        return plan

    def _copy_in(self, plan, image):
        if isinstance(image, torch.Tensor):
            plan.image.copy_(image, non_blocking=True)
        else:
            plan.image.copy_(torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)), non_blocking=True)

    def _run(self, image, im_info, post=False, detect=False, orig_hw=None):
        assert image.shape[0] == 1 and image.shape[3] == 3, "test_image / im_detect take ONE image (use detect_batch for more)"
        plan = self.plan_for(image.shape[1], image.shape[2])
        self._copy_in(plan, image)
        oh, ow = orig_hw if orig_hw is not None else (None, None)
        plan.launch(float(im_info[2]), oh, ow, post=post, detect=detect)
        return plan

    def detect(self, image, im_info, orig_hw):
        """Fused device path for im_detect + test_net's per-class NMS + max_per_image cap.
        Returns (det [n,6] = x1,y1,x2,y2,score,class; plan) after one stream sync (one D2H copy of the record)."""
        plan = self._run(image, im_info, post=True, detect=True, orig_hw=orig_hw)
        return plan.records()[0], plan

    def detect_batch(self, images, im_scales, orig_hws):
        """Throughput path: `images` [B,H,W,3] blobs of ONE shape (numpy or a pinned torch tensor), per-image scale factors and
        original (h, w).  -> (list of B det arrays [n,6], plan).  The reference is batch 1; this is SURVEY 8(f) rank 4."""
        b = int(images.shape[0])
        assert images.shape[3] == 3 and len(im_scales) == b and len(orig_hws) == b
        plan = self.plan_for(images.shape[1], images.shape[2], b)
        self._copy_in(plan, images)
        plan.launch(post=True, detect=True, meta=[(float(im_scales[i]), int(orig_hws[i][0]), int(orig_hws[i][1])) for i in range(b)])
        return plan.records(), plan

    # ---- pipelined throughput API: overlap the next batch's host->device copy with the current batch's compute ---------------
    def submit_batch(self, images, im_scales, orig_hws):
        """Enqueue one batch without waiting for it: the H2D copy of `images` (a PINNED torch tensor or a numpy array, [B,H,W,3])
        runs on a copy stream into one of two staging buffers, the compute stream picks it up behind an event, replays the graph
        and copies the records into one of two pinned host buffers.  -> ticket for collect_batch().  At most TWO tickets may be
        outstanding (submit i+1, collect i, submit i+2, ...): that is what keeps the copy of batch i+1 under the compute of
        batch i.  Results are identical to detect_batch()."""
        b = int(images.shape[0])
        assert images.shape[3] == 3 and len(im_scales) == b and len(orig_hws) == b
        plan = self.plan_for(images.shape[1], images.shape[2], b)
        pipe = plan.__dict__.setdefault("_pipe", None)
        if pipe is None:
            pipe = plan._pipe = dict(copy_stream=torch.cuda.Stream(), stage=[torch.empty_like(plan.image) for _ in range(2)],
                                     staged=[torch.cuda.Event() for _ in range(2)], consumed=[torch.cuda.Event() for _ in range(2)],
                                     done=[torch.cuda.Event() for _ in range(2)], host=[None, None], turn=0, used=[False, False])
        k = pipe["turn"] & 1
        pipe["turn"] += 1
        src = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32))
        main = torch.cuda.current_stream()
        with torch.cuda.stream(pipe["copy_stream"]):
            if pipe["used"][k]:
                pipe["copy_stream"].wait_event(pipe["consumed"][k])     # the compute stream has taken the previous content of stage k
            pipe["stage"][k].copy_(src, non_blocking=True)
            pipe["staged"][k].record(pipe["copy_stream"])
        main.wait_event(pipe["staged"][k])
        plan.image.copy_(pipe["stage"][k], non_blocking=True)             # device-to-device: ~microseconds, keeps the graph's input address fixed
        pipe["consumed"][k].record(main)
        pipe["used"][k] = True
        plan.launch(post=True, detect=True, meta=[(float(im_scales[i]), int(orig_hws[i][0]), int(orig_hws[i][1])) for i in range(b)])
        if pipe["host"][k] is None or pipe["host"][k].shape != plan.rec.shape:
            pipe["host"][k] = torch.empty(plan.rec.shape, dtype=torch.float32).pin_memory()
        pipe["host"][k].copy_(plan.rec, non_blocking=True)
        pipe["done"][k].record(main)
        return (plan, k, plan.max_det)

    def collect_batch(self, ticket):
        """Wait for a submitted batch -> list of B det arrays [n,6]."""
        plan, k, max_det = ticket
        pipe = plan._pipe
        pipe["done"][k].synchronize()
        return engine.split_host_records(pipe["host"][k], max_det)
