"""PASCAL VOC image database for the TEST path (lib/datasets/pascal_voc.py:25-290): image-set index, image paths,
ground-truth roidb, result files in the devkit layout, Python AP evaluation, competition mode.

    <devkit>/VOC<year>/ImageSets/Main/<split>.txt      one image id per line
    <devkit>/VOC<year>/JPEGImages/<id>.jpg
    <devkit>/VOC<year>/Annotations/<id>.xml
    <devkit>/results/VOC<year>/Main/<comp_id>_det_<split>_<class>.txt   written by evaluate_detections

`<devkit>` defaults to <cfg.DATA_DIR>/VOCdevkit<year> as in the reference; `devkit_path=` overrides it.  The MATLAB
evaluation and the selective-search / RPN proposal roidbs (training inputs) are not provided."""
import os
import pickle
import uuid
import xml.etree.ElementTree as ET

import numpy as np
import scipy.sparse

from datasets.imdb import imdb
from datasets.voc_eval import voc_eval
from model.config import cfg

VOC_CLASSES = ("__background__",
               "aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable",
               "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor")


class pascal_voc(imdb):
    def __init__(self, image_set, year, use_diff=False, devkit_path=None):
        imdb.__init__(self, "voc_" + year + "_" + image_set + ("_diff" if use_diff else ""), VOC_CLASSES)
        self._year = year
        self._image_set = image_set
        self._devkit_path = devkit_path or os.path.join(cfg.DATA_DIR, "VOCdevkit" + year)
        self._data_path = os.path.join(self._devkit_path, "VOC" + year)
        for path, what in ((self._devkit_path, "VOCdevkit path"), (self._data_path, "Path")):
            if not os.path.exists(path):
                raise AssertionError("{} does not exist: {}".format(what, path))
        self._class_to_ind = {c: i for i, c in enumerate(self._classes)}
        self._image_ext = ".jpg"
        self._image_index = self._load_image_set_index()
        self._salt = str(uuid.uuid4())
        self._comp_id = "comp4"
        self.config = {"cleanup": True, "use_salt": True, "use_diff": use_diff, "matlab_eval": False, "rpn_file": None}

    # ---- index / paths -------------------------------------------------------------------------------------------
    def _in_data(self, *parts):
        return os.path.join(self._data_path, *parts)

    def _load_image_set_index(self):
        listing = self._in_data("ImageSets", "Main", self._image_set + ".txt")
        if not os.path.exists(listing):
            raise AssertionError("Path does not exist: {}".format(listing))
        with open(listing) as f:
            return [line.strip() for line in f if line.strip()]

    def image_path_from_index(self, index):
        path = self._in_data("JPEGImages", index + self._image_ext)
        if not os.path.exists(path):
            raise AssertionError("Path does not exist: {}".format(path))
        return path

    def image_path_at(self, i):
        return self.image_path_from_index(self._image_index[i])

    # ---- ground truth ----------------------------------------------------------------------------------------------
    def _load_pascal_annotation(self, index):
        """Boxes become 0-based (xml is 1-based) and are stored as uint16, 'difficult' objects are dropped unless
        use_diff, seg_areas is the inclusive box area (pascal_voc.py:140-183)."""
        objs = ET.parse(self._in_data("Annotations", index + ".xml")).findall("object")
        if not self.config["use_diff"]:
            objs = [o for o in objs if int(o.find("difficult").text) == 0]
        n = len(objs)
        boxes = np.zeros((n, 4), dtype=np.uint16)
        gt_classes = np.zeros(n, dtype=np.int32)
        overlaps = np.zeros((n, self.num_classes), dtype=np.float32)
        seg_areas = np.zeros(n, dtype=np.float32)
        for k, obj in enumerate(objs):
            bb = obj.find("bndbox")
            x1, y1, x2, y2 = (float(bb.find(t).text) - 1 for t in ("xmin", "ymin", "xmax", "ymax"))
            cls = self._class_to_ind[obj.find("name").text.lower().strip()]
            boxes[k] = (x1, y1, x2, y2)
            gt_classes[k] = cls
            overlaps[k, cls] = 1.0
            seg_areas[k] = (x2 - x1 + 1) * (y2 - y1 + 1)
        return {"boxes": boxes, "gt_classes": gt_classes, "gt_overlaps": scipy.sparse.csr_matrix(overlaps),
                "flipped": False, "seg_areas": seg_areas}

    def gt_roidb(self):
        cache_file = os.path.join(self.cache_path, self.name + "_gt_roidb.pkl")
        if os.path.exists(cache_file):
            with open(cache_file, "rb") as f:
                roidb = pickle.load(f)
            print("{} gt roidb loaded from {}".format(self.name, cache_file))
            return roidb
        roidb = [self._load_pascal_annotation(index) for index in self._image_index]
        with open(cache_file, "wb") as f:
            pickle.dump(roidb, f, pickle.HIGHEST_PROTOCOL)
        print("wrote gt roidb to {}".format(cache_file))
        return roidb

    # ---- results + evaluation ----------------------------------------------------------------------------------------
    def _get_comp_id(self):
        return self._comp_id + "_" + self._salt if self.config["use_salt"] else self._comp_id

    def _get_voc_results_file_template(self):
        name = self._get_comp_id() + "_det_" + self._image_set + "_{:s}.txt"
        return os.path.join(self._devkit_path, "results", "VOC" + self._year, "Main", name)

    def _write_voc_results_file(self, all_boxes):
        """`<image id> <score %.3f> <x1+1 %.1f> <y1+1> <x2+1> <y2+1>` per detection (the devkit is 1-based)."""
        template = self._get_voc_results_file_template()
        os.makedirs(os.path.dirname(template), exist_ok=True)
        for c, cls in enumerate(self._classes):
            if c == 0:
                continue
            print("Writing {} VOC results file".format(cls))
            with open(template.format(cls), "wt") as f:
                for i, index in enumerate(self._image_index):
                    dets = np.asarray(all_boxes[c][i])
                    if dets.size == 0:
                        continue
                    for d in dets.reshape(-1, dets.shape[-1]):
                        f.write("{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n".format(
                            index, d[-1], d[0] + 1, d[1] + 1, d[2] + 1, d[3] + 1))

    def _do_python_eval(self, output_dir="output"):
        annopath = self._in_data("Annotations", "{:s}.xml")
        imagesetfile = self._in_data("ImageSets", "Main", self._image_set + ".txt")
        cachedir = os.path.join(self._devkit_path, "annotations_cache")
        use_07_metric = int(self._year) < 2010            # the VOC metric changed in 2010
        print("VOC07 metric? " + ("Yes" if use_07_metric else "No"))
        os.makedirs(output_dir, exist_ok=True)
        aps = []
        for cls in self._classes[1:]:
            rec, prec, ap = voc_eval(self._get_voc_results_file_template().format(cls), annopath, imagesetfile, cls,
                                     cachedir, ovthresh=0.5, use_07_metric=use_07_metric, use_diff=self.config["use_diff"])
            aps.append(ap)
            print("AP for {} = {:.4f}".format(cls, ap))
            with open(os.path.join(output_dir, cls + "_pr.pkl"), "wb") as f:
                pickle.dump({"rec": rec, "prec": prec, "ap": ap}, f)
        print("Mean AP = {:.4f}".format(np.mean(aps)))
        print("~~~~~~~~\nResults:")
        for ap in aps:
            print("{:.3f}".format(ap))
        print("{:.3f}\n~~~~~~~~".format(np.mean(aps)))
        print("Computed with the unofficial Python evaluation (see the reference's note: use the MATLAB devkit for papers).")
        return aps

    def evaluate_detections(self, all_boxes, output_dir=None):
        self._write_voc_results_file(all_boxes)
        aps = self._do_python_eval(output_dir or "output")
        if self.config["matlab_eval"]:
            raise NotImplementedError("MATLAB evaluation is not provided; use the devkit on the kept result files "
                                      "(competition_mode(True) keeps them)")
        if self.config["cleanup"]:
            for cls in self._classes[1:]:
                os.remove(self._get_voc_results_file_template().format(cls))
        return aps

    def competition_mode(self, on):
        self.config["use_salt"] = not on
        self.config["cleanup"] = not on
