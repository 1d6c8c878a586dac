"""COCO bounding-box detection metrics without pycocotools (absent from this image).  The reference scores COCO runs
with `pycocotools.cocoeval.COCOeval` (lib/datasets/coco.py:246-256) -- third-party code that is not under
/root/reference; this module restates its published bbox protocol:

  * IoU thresholds 0.50:0.05:0.95, 101 recall points, area ranges all / small (<32^2) / medium / large (>96^2),
    at most 1 / 10 / 100 detections per image, per category;
  * box IoU on [x, y, w, h] without the '+1' convention; against a crowd region the denominator is the detection area;
  * per image and category: detections in descending score order (stable) are greedily matched to the not yet matched
    ground-truth box of highest IoU >= threshold, preferring non-ignored boxes; crowd boxes may be matched repeatedly;
    ground truth that is crowd or outside the area range is 'ignored', and so are detections matched to it or unmatched
    detections outside the range;
  * precision is made monotone from the right and sampled at the recall points (first index with recall >= point),
    -1 marks category/area cells without ground truth; AP averages all cells > -1.

**Parity unpinned**: no pycocotools here to compare with; tests/test_datasets.py holds hand-computed cases only."""
import numpy as np

IOU_THRS = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1, endpoint=True)
REC_THRS = np.linspace(0.0, 1.0, int(np.round((1.0 - 0.0) / 0.01)) + 1, endpoint=True)
MAX_DETS = (1, 10, 100)
AREA_RNG = ((0.0, 1e10), (0.0, 32.0 ** 2), (32.0 ** 2, 96.0 ** 2), (96.0 ** 2, 1e10))
AREA_LBL = ("all", "small", "medium", "large")


def box_iou(dt, gt, iscrowd):
    """dt [D,4], gt [G,4] as x, y, w, h -> [D,G]."""
    dt = np.asarray(dt, dtype=np.float64).reshape(-1, 4)
    gt = np.asarray(gt, dtype=np.float64).reshape(-1, 4)
    crowd = np.asarray(iscrowd, dtype=bool).reshape(-1)
    w = np.minimum(dt[:, None, 0] + dt[:, None, 2], gt[None, :, 0] + gt[None, :, 2]) - np.maximum(dt[:, None, 0], gt[None, :, 0])
    h = np.minimum(dt[:, None, 1] + dt[:, None, 3], gt[None, :, 1] + gt[None, :, 3]) - np.maximum(dt[:, None, 1], gt[None, :, 1])
    inter = np.clip(w, 0, None) * np.clip(h, 0, None)
    da = (dt[:, 2] * dt[:, 3])[:, None]
    ga = (gt[:, 2] * gt[:, 3])[None, :]
    union = np.where(crowd[None, :], da, da + ga - inter)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(union > 0, inter / union, 0.0)
    return out


def _match_image(dt, gt, area_rng, max_det):
    """dt: (boxes [D,4], scores [D]) already sorted by score, descending; gt: (boxes [G,4], areas [G], crowd [G]).
    Returns (scores [d], matched [T,d] bool, dt_ignored [T,d] bool, gt_ignored [G] bool) or None if both are empty."""
    d_boxes, d_scores = dt
    g_boxes, g_areas, g_crowd = gt
    if len(d_scores) == 0 and len(g_areas) == 0:
        return None
    g_ign = g_crowd | (g_areas < area_rng[0]) | (g_areas > area_rng[1])
    g_order = np.argsort(g_ign, kind="mergesort")               # non-ignored first, stable
    g_boxes, g_crowd, g_ign = g_boxes[g_order], g_crowd[g_order], g_ign[g_order]
    d_boxes, d_scores = d_boxes[:max_det], d_scores[:max_det]
    ious = box_iou(d_boxes, g_boxes, g_crowd)
    T, D, G = len(IOU_THRS), len(d_scores), len(g_ign)
    g_taken = np.zeros((T, G), dtype=bool)
    d_match = np.zeros((T, D), dtype=bool)
    d_ign = np.zeros((T, D), dtype=bool)
    for t, thr in enumerate(IOU_THRS):
        for d in range(D):
            best = min(thr, 1 - 1e-10)
            m = -1
            for g in range(G):
                if g_taken[t, g] and not g_crowd[g]:
                    continue
                if m > -1 and not g_ign[m] and g_ign[g]:
                    break                                      # a real match exists; only ignored boxes follow
                if ious[d, g] < best:
                    continue
                best = ious[d, g]
                m = g
            if m == -1:
                continue
            d_ign[t, d] = g_ign[m]
            d_match[t, d] = True
            g_taken[t, m] = True
    d_area = d_boxes[:, 2] * d_boxes[:, 3]
    outside = (d_area < area_rng[0]) | (d_area > area_rng[1])
    d_ign |= (~d_match) & outside[None, :]
    return d_scores, d_match, d_ign, g_ign


class BboxEval(object):
    """evaluate() + accumulate() + summarize() over
       gt: {image_id: [{'bbox': [x,y,w,h], 'area': a, 'iscrowd': 0/1, 'category_id': c}]},
       dt: [{'image_id', 'category_id', 'bbox': [x,y,w,h], 'score'}],  cat_ids / img_ids: evaluation order."""

    def __init__(self, gt, dt, cat_ids, img_ids):
        self.cat_ids = list(cat_ids)
        self.img_ids = list(img_ids)
        self._gt = {}
        self._dt = {}
        for img, anns in gt.items():
            for a in anns:
                self._gt.setdefault((img, a["category_id"]), []).append(a)
        for d in dt:
            self._dt.setdefault((d["image_id"], d["category_id"]), []).append(d)
        self.precision = None
        self.recall = None
        self.stats = None

    def _pair(self, img, cat):
        g = self._gt.get((img, cat), [])
        d = self._dt.get((img, cat), [])
        scores = np.array([x["score"] for x in d], dtype=np.float64)
        order = np.argsort(-scores, kind="mergesort")[:MAX_DETS[-1]]
        d_boxes = np.array([d[i]["bbox"] for i in order], dtype=np.float64).reshape(-1, 4)
        g_boxes = np.array([x["bbox"] for x in g], dtype=np.float64).reshape(-1, 4)
        g_areas = np.array([x["area"] for x in g], dtype=np.float64)
        g_crowd = np.array([bool(x.get("iscrowd", 0)) for x in g], dtype=bool)
        return (d_boxes, scores[order]), (g_boxes, g_areas, g_crowd)

    def evaluate(self):
        K, A, M = len(self.cat_ids), len(AREA_RNG), len(MAX_DETS)
        T, R = len(IOU_THRS), len(REC_THRS)
        self.precision = -np.ones((T, R, K, A, M))
        self.recall = -np.ones((T, K, A, M))
        for k, cat in enumerate(self.cat_ids):
            pairs = [self._pair(img, cat) for img in self.img_ids]
            for a, rng in enumerate(AREA_RNG):
                for m, max_det in enumerate(MAX_DETS):
                    per_img = [r for r in (_match_image(dt, gt, rng, max_det) for dt, gt in pairs) if r is not None]
                    if not per_img:
                        continue
                    self._accumulate_cell(per_img, k, a, m)
        return self

    def _accumulate_cell(self, per_img, k, a, m):
        scores = np.concatenate([r[0] for r in per_img])
        order = np.argsort(-scores, kind="mergesort")
        matched = np.concatenate([r[1] for r in per_img], axis=1)[:, order]
        ignored = np.concatenate([r[2] for r in per_img], axis=1)[:, order]
        npig = int(np.count_nonzero(~np.concatenate([r[3] for r in per_img])))
        if npig == 0:
            return
        tp_sum = np.cumsum(matched & ~ignored, axis=1).astype(np.float64)
        fp_sum = np.cumsum(~matched & ~ignored, axis=1).astype(np.float64)
        for t in range(len(IOU_THRS)):
            tp, fp = tp_sum[t], fp_sum[t]
            nd = len(tp)
            rc = tp / npig
            pr = tp / (fp + tp + np.spacing(1))
            self.recall[t, k, a, m] = rc[-1] if nd else 0.0
            pr = np.maximum.accumulate(pr[::-1])[::-1] if nd else pr
            idx = np.searchsorted(rc, REC_THRS, side="left")
            q = np.zeros(len(REC_THRS))
            ok = idx < nd
            q[ok] = pr[idx[ok]]
            self.precision[t, :, k, a, m] = q

    accumulate = lambda self: self                                # evaluate() already fills the tables

    def _stat(self, ap, iou_thr=None, area="all", max_det=100):
        a = AREA_LBL.index(area)
        m = MAX_DETS.index(max_det)
        s = (self.precision[:, :, :, a, m] if ap else self.recall[:, :, a, m])
        if iou_thr is not None:
            s = s[np.where(np.isclose(IOU_THRS, iou_thr))[0]]
        s = s[s > -1]
        return float(np.mean(s)) if s.size else -1.0

    def summarize(self, verbose=True):
        spec = [(1, None, "all", 100), (1, 0.5, "all", 100), (1, 0.75, "all", 100), (1, None, "small", 100),
                (1, None, "medium", 100), (1, None, "large", 100), (0, None, "all", 1), (0, None, "all", 10),
                (0, None, "all", 100), (0, None, "small", 100), (0, None, "medium", 100), (0, None, "large", 100)]
        self.stats = np.array([self._stat(*s) for s in spec])
        if verbose:
            for (ap, thr, area, md), v in zip(spec, self.stats):
                print(" {:<18} {} @[ IoU={:<9} | area={:>6s} | maxDets={:>3d} ] = {:0.3f}".format(
                    "Average Precision" if ap else "Average Recall", "(AP)" if ap else "(AR)",
                    "0.50:0.95" if thr is None else "{:0.2f}".format(thr), area, md, v))
        return self.stats
