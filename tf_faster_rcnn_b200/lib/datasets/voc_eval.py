"""PASCAL VOC detection scoring, the consumer of test_net's result files (lib/datasets/voc_eval.py:15-214):
`parse_rec`, `voc_ap`, `voc_eval` keep the reference's signatures and return values.  Differences: the annotation
cache is written in binary mode (the reference opens it with 'w', which fails under Python 3), and matching runs on
per-image arrays prepared once instead of re-deriving them per detection; tests/test_datasets.py checks rec/prec/ap
against values produced by the reference's own function (tests/golden/make_golden.py)."""
import os
import pickle
import xml.etree.ElementTree as ET

import numpy as np


def parse_rec(filename):
    """One VOC annotation xml -> [{'name', 'pose', 'truncated', 'difficult', 'bbox': [xmin, ymin, xmax, ymax]}]
    (1-based pixel coordinates, as stored)."""
    objects = []
    for node in ET.parse(filename).findall("object"):
        box = node.find("bndbox")
        objects.append({
            "name": node.find("name").text,
            "pose": node.find("pose").text,
            "truncated": int(node.find("truncated").text),
            "difficult": int(node.find("difficult").text),
            "bbox": [int(box.find(tag).text) for tag in ("xmin", "ymin", "xmax", "ymax")],
        })
    return objects


def voc_ap(rec, prec, use_07_metric=False):
    """Average precision from cumulative recall / precision.  VOC07: mean over recall levels 0, 0.1 .. 1 of the best
    precision at recall >= level.  Otherwise: area under the monotone (right-to-left running max) precision envelope."""
    rec = np.asarray(rec, dtype=np.float64)
    prec = np.asarray(prec, dtype=np.float64)
    if use_07_metric:
        ap = 0.0
        for level in np.arange(0.0, 1.1, 0.1):
            reached = rec >= level
            ap = ap + (np.max(prec[reached]) if reached.any() else 0.0) / 11.0
        return ap
    r = np.concatenate(([0.0], rec, [1.0]))
    p = np.concatenate(([0.0], prec, [0.0]))
    p = np.maximum.accumulate(p[::-1])[::-1]
    steps = np.nonzero(r[1:] != r[:-1])[0]
    return np.sum((r[steps + 1] - r[steps]) * p[steps + 1])


def _load_annotations(annopath, imagenames, cachefile):
    if os.path.isfile(cachefile):
        with open(cachefile, "rb") as f:
            try:
                return pickle.load(f)
            except UnicodeDecodeError:
                f.seek(0)
                return pickle.load(f, encoding="bytes")
    recs = {}
    for i, name in enumerate(imagenames):
        recs[name] = parse_rec(annopath.format(name))
        if i % 100 == 0:
            print("Reading annotation for {:d}/{:d}".format(i + 1, len(imagenames)))
    print("Saving cached annotations to {:s}".format(cachefile))
    with open(cachefile, "wb") as f:
        pickle.dump(recs, f)
    return recs


def voc_eval(detpath, annopath, imagesetfile, classname, cachedir, ovthresh=0.5, use_07_metric=False, use_diff=False):
    """rec, prec, ap for one class.  detpath.format(classname): result file of `<image id> <score> <x1> <y1> <x2> <y2>`
    lines (1-based); annopath.format(image id): annotation xml; imagesetfile: one image id per line.
    A detection is a true positive when its best-overlapping ground-truth box of the class (inclusive '+1' IoU,
    strictly greater than ovthresh) is not 'difficult' and not yet claimed by a higher-scoring detection; a second claim
    is a false positive; a hit on a 'difficult' box is ignored (unless use_diff)."""
    os.makedirs(cachedir, exist_ok=True)
    with open(imagesetfile, "r") as f:
        imagenames = [line.strip() for line in f]
    # the reference names the cache after the *full path* of the image-set file under cachedir (os.path.join with an
    # absolute second part yields "<imagesetfile>_annots.pkl" beside the image-set file); keep that location
    recs = _load_annotations(annopath, imagenames, os.path.join(cachedir, "%s_annots.pkl" % imagesetfile))

    gt = {}
    npos = 0
    for name in imagenames:
        objs = [o for o in recs[name] if o["name"] == classname]
        boxes = np.array([o["bbox"] for o in objs], dtype=np.float64).reshape(-1, 4)
        difficult = np.zeros(len(objs), dtype=bool) if use_diff else np.array([o["difficult"] for o in objs], dtype=bool)
        npos += int((~difficult).sum())
        areas = (boxes[:, 2] - boxes[:, 0] + 1.0) * (boxes[:, 3] - boxes[:, 1] + 1.0)
        gt[name] = (boxes, difficult, np.zeros(len(objs), dtype=bool), areas)

    with open(detpath.format(classname), "r") as f:
        rows = [line.strip().split(" ") for line in f if line.strip()]
    ids = [r[0] for r in rows]
    conf = np.array([float(r[1]) for r in rows], dtype=np.float64)
    dets = np.array([[float(v) for v in r[2:6]] for r in rows], dtype=np.float64).reshape(-1, 4)

    nd = len(ids)
    tp = np.zeros(nd)
    fp = np.zeros(nd)
    order = np.argsort(-conf)
    for rank, d in enumerate(order):
        boxes, difficult, claimed, areas = gt[ids[d]]
        bb = dets[d]
        best = -np.inf
        if boxes.shape[0]:
            iw = np.maximum(np.minimum(boxes[:, 2], bb[2]) - np.maximum(boxes[:, 0], bb[0]) + 1.0, 0.0)
            ih = np.maximum(np.minimum(boxes[:, 3], bb[3]) - np.maximum(boxes[:, 1], bb[1]) + 1.0, 0.0)
            inter = iw * ih
            iou = inter / ((bb[2] - bb[0] + 1.0) * (bb[3] - bb[1] + 1.0) + areas - inter)
            j = int(np.argmax(iou))
            best = iou[j]
        if best > ovthresh:
            if difficult[j]:
                continue
            if claimed[j]:
                fp[rank] = 1.0
            else:
                tp[rank] = 1.0
                claimed[j] = True
        else:
            fp[rank] = 1.0

    fp = np.cumsum(fp)
    tp = np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)
