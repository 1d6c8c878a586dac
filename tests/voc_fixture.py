"""Deterministic miniature VOCdevkit tree + detections, shared by tests/golden/make_voc_golden.py (which feeds it to
the reference's voc_eval) and tests/test_datasets.py (which feeds it to this repo's)."""
import os

import numpy as np

CLASSES = ("aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable", "dog",
           "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor")
USED = ("car", "person", "dog", "boat")          # classes that receive objects / detections; "boat": gt but no detection file rows


def _xml(objs, w, h):
    body = "".join(
        "<object><name>%s</name><pose>Unspecified</pose><truncated>%d</truncated><difficult>%d</difficult>"
        "<bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax><ymax>%d</ymax></bndbox></object>" % o for o in objs)
    return "<annotation><size><width>%d</width><height>%d</height><depth>3</depth></size>%s</annotation>" % (w, h, body)


def build(root, year="2007", split="test", n_images=24, seed=7, with_images=False):
    """Writes <root>/VOC<year>/{ImageSets/Main/<split>.txt, Annotations/*.xml[, JPEGImages/*.jpg]} and returns
    (image ids, {class: [(image id, score, x1, y1, x2, y2)]}) -- detection rows already in devkit (1-based) coordinates."""
    rng = np.random.default_rng(seed)
    pixel_rng = np.random.default_rng(seed + 1)          # separate stream: annotations do not depend on with_images
    data = os.path.join(root, "VOC" + year)
    for sub in ("ImageSets/Main", "Annotations", "JPEGImages"):
        os.makedirs(os.path.join(data, sub), exist_ok=True)
    ids = ["%06d" % (i * 3 + 1) for i in range(n_images)]
    with open(os.path.join(data, "ImageSets", "Main", split + ".txt"), "w") as f:
        f.write("".join(i + "\n" for i in ids))
    dets = {c: [] for c in CLASSES}
    for img in ids:
        w, h = int(rng.integers(200, 500)), int(rng.integers(200, 400))
        objs = []
        for _ in range(int(rng.integers(0, 6))):
            cls = USED[int(rng.integers(0, len(USED)))]
            x1, y1 = int(rng.integers(1, w - 60)), int(rng.integers(1, h - 60))
            x2, y2 = int(rng.integers(x1 + 20, w)), int(rng.integers(y1 + 20, h))
            difficult = int(rng.random() < 0.25)
            objs.append((cls, int(rng.random() < 0.3), difficult, x1, y1, x2, y2))
            if cls == "boat":
                continue
            # 0-2 detections near the object (some duplicates -> double-claim false positives), jittered
            for _ in range(int(rng.integers(0, 3))):
                j = rng.normal(0, 12, 4)
                dets[cls].append((img, float(rng.random()), x1 + j[0], y1 + j[1], x2 + j[2], y2 + j[3]))
        for cls in ("car", "person", "dog"):               # background false positives
            for _ in range(int(rng.integers(0, 2))):
                x1, y1 = rng.uniform(1, w - 50), rng.uniform(1, h - 50)
                dets[cls].append((img, float(rng.random()) * 0.8, x1, y1, x1 + rng.uniform(10, 49), y1 + rng.uniform(10, 49)))
        with open(os.path.join(data, "Annotations", img + ".xml"), "w") as f:
            f.write(_xml(objs, w, h))
        if with_images:
            import cv2
            cv2.imwrite(os.path.join(data, "JPEGImages", img + ".jpg"), pixel_rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    return ids, dets


def write_det_files(dirname, dets, pattern="det_{:s}.txt"):
    os.makedirs(dirname, exist_ok=True)
    for cls, rows in dets.items():
        with open(os.path.join(dirname, pattern.format(cls)), "w") as f:
            for r in rows:
                f.write("{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n".format(*r))
    return os.path.join(dirname, pattern)
