"""MS-COCO image database for the TEST path (lib/datasets/coco.py:27-316): annotation index, image paths, ground-truth
roidb with the reference's box sanitising, results json, bbox evaluation, competition mode.

    <DATA_DIR>/coco/annotations/{instances,image_info}_<split><year>.json
    <DATA_DIR>/coco/images/<data split>/COCO_<data split>_<12-digit id>.jpg

pycocotools is not required: the annotation json is indexed here directly (same orders as the COCO API: images and
categories in file order) and detections are scored by datasets/coco_eval.py unless pycocotools is importable."""
import json
import os
import pickle
import uuid

import numpy as np
import scipy.sparse

from datasets.imdb import imdb
from model.config import cfg

_VIEW = {"minival2014": "val2014", "valminusminival2014": "val2014", "test-dev2015": "test2015"}


class _AnnotationIndex(object):
    def __init__(self, path):
        with open(path) as f:
            data = json.load(f)
        self.images = {im["id"]: im for im in data.get("images", [])}
        self.categories = list(data.get("categories", []))
        self.by_image = {}
        for ann in data.get("annotations", []):
            self.by_image.setdefault(ann["image_id"], []).append(ann)


class coco(imdb):
    def __init__(self, image_set, year, data_path=None):
        imdb.__init__(self, "coco_" + year + "_" + image_set)
        self.config = {"use_salt": True, "cleanup": True}
        self._year = year
        self._image_set = image_set
        self._data_path = data_path or os.path.join(cfg.DATA_DIR, "coco")
        self._index = _AnnotationIndex(self._get_ann_file())
        cats = self._index.categories
        self._classes = tuple(["__background__"] + [c["name"] for c in cats])
        self._class_to_ind = {c: i for i, c in enumerate(self._classes)}
        self._class_to_coco_cat_id = {c["name"]: c["id"] for c in cats}
        self._image_index = list(self._index.images.keys())
        name = image_set + year
        self._data_name = _VIEW.get(name, name)
        self._gt_splits = ("train", "val", "minival")

    def _get_ann_file(self):
        prefix = "instances" if "test" not in self._image_set else "image_info"
        return os.path.join(self._data_path, "annotations", prefix + "_" + self._image_set + self._year + ".json")

    def image_path_from_index(self, index):
        path = os.path.join(self._data_path, "images", self._data_name,
                            "COCO_" + self._data_name + "_" + str(index).zfill(12) + ".jpg")
        if not os.path.exists(path):
            raise AssertionError("Path does not exist: {}".format(path))
        return path

    def image_path_at(self, i):
        return self.image_path_from_index(self._image_index[i])

    # ---- ground truth ----------------------------------------------------------------------------------------------
    def _load_coco_annotation(self, index):
        """x2 = min(W-1, x1 + max(0, w-1)) etc.; zero-area or inverted boxes dropped; crowd rows get overlap -1 for every
        class (coco.py:122-178)."""
        im = self._index.images[index]
        width, height = im["width"], im["height"]
        rows = []
        for obj in self._index.by_image.get(index, []):
            x1 = max(0, obj["bbox"][0])
            y1 = max(0, obj["bbox"][1])
            x2 = min(width - 1, x1 + max(0, obj["bbox"][2] - 1))
            y2 = min(height - 1, y1 + max(0, obj["bbox"][3] - 1))
            if obj["area"] > 0 and x2 >= x1 and y2 >= y1:
                rows.append((obj, (x1, y1, x2, y2)))
        n = len(rows)
        boxes = np.zeros((n, 4), dtype=np.uint16)
        gt_classes = np.zeros(n, dtype=np.int32)
        overlaps = np.zeros((n, self.num_classes), dtype=np.float32)
        seg_areas = np.zeros(n, dtype=np.float32)
        cat_to_ind = {self._class_to_coco_cat_id[c]: self._class_to_ind[c] for c in self._classes[1:]}
        for k, (obj, box) in enumerate(rows):
            cls = cat_to_ind[obj["category_id"]]
            boxes[k] = box
            gt_classes[k] = cls
            seg_areas[k] = obj["area"]
            if obj.get("iscrowd", 0):
                overlaps[k, :] = -1.0
            else:
                overlaps[k, cls] = 1.0
        assert (boxes[:, 2] >= boxes[:, 0]).all() and (boxes[:, 3] >= boxes[:, 1]).all()
        assert (boxes[:, 2] < width).all() and (boxes[:, 3] < height).all()
        return {"width": width, "height": height, "boxes": boxes, "gt_classes": gt_classes,
                "gt_overlaps": scipy.sparse.csr_matrix(overlaps), "flipped": False, "seg_areas": seg_areas}

    def gt_roidb(self):
        cache_file = os.path.join(self.cache_path, self.name + "_gt_roidb.pkl")
        if os.path.exists(cache_file):
            with open(cache_file, "rb") as f:
                roidb = pickle.load(f)
            print("{} gt roidb loaded from {}".format(self.name, cache_file))
            return roidb
        roidb = [self._load_coco_annotation(index) for index in self._image_index]
        with open(cache_file, "wb") as f:
            pickle.dump(roidb, f, pickle.HIGHEST_PROTOCOL)
        print("wrote gt roidb to {}".format(cache_file))
        return roidb

    # ---- results + evaluation ----------------------------------------------------------------------------------------
    def _coco_results_one_category(self, boxes, cat_id):
        out = []
        for i, index in enumerate(self._image_index):
            dets = np.asarray(boxes[i], dtype=np.float64)
            if dets.size == 0:
                continue
            for x1, y1, x2, y2, score in dets.reshape(-1, dets.shape[-1])[:, :5]:
                out.append({"image_id": index, "category_id": cat_id,
                            "bbox": [float(x1), float(y1), float(x2 - x1 + 1), float(y2 - y1 + 1)], "score": float(score)})
        return out

    def _write_coco_results_file(self, all_boxes, res_file):
        results = []
        for c, cls in enumerate(self._classes):
            if c == 0:
                continue
            print("Collecting {} results ({:d}/{:d})".format(cls, c, self.num_classes - 1))
            results.extend(self._coco_results_one_category(all_boxes[c], self._class_to_coco_cat_id[cls]))
        print("Writing results json to {}".format(res_file))
        with open(res_file, "w") as f:
            json.dump(results, f)
        return results

    def _print_detection_eval_metrics(self, precision):
        """precision [T, R, K, A, M]: mean and per-category AP over IoU 0.50:0.95, all areas, 100 detections (coco.py:207-241)."""
        p = precision[:, :, :, 0, 2]
        print("~~~~ Mean and per-category AP @ IoU=[0.50,0.95] ~~~~")
        valid = p[p > -1]
        mean_ap = float(np.mean(valid)) if valid.size else float("nan")
        print("{:.1f}".format(100 * mean_ap))
        per_cat = []
        for k in range(p.shape[2]):
            v = p[:, :, k]
            v = v[v > -1]
            per_cat.append(float(np.mean(v)) if v.size else float("nan"))
            print("{:.1f}".format(100 * per_cat[-1]))
        return mean_ap, per_cat

    def _do_detection_eval(self, res_file, output_dir):
        try:
            from pycocotools.coco import COCO
            from pycocotools.cocoeval import COCOeval
        except ImportError:
            COCO = None
        if COCO is not None:
            gt_api = COCO(self._get_ann_file())
            ev = COCOeval(gt_api, gt_api.loadRes(res_file), "bbox")
            ev.evaluate()
            ev.accumulate()
            precision = ev.eval["precision"]
            mean_ap, _ = self._print_detection_eval_metrics(precision)
            print("~~~~ Summary metrics ~~~~")
            ev.summarize()
            stats = np.asarray(ev.stats)
        else:
            from datasets.coco_eval import BboxEval
            with open(res_file) as f:
                dt = json.load(f)
            cat_ids = [c["id"] for c in self._index.categories]
            ev = BboxEval(self._index.by_image, dt, cat_ids, self._image_index).evaluate()
            precision = ev.precision
            mean_ap, _ = self._print_detection_eval_metrics(precision)
            print("~~~~ Summary metrics ~~~~")
            stats = ev.summarize()
        eval_file = os.path.join(output_dir, "detection_results.pkl")
        with open(eval_file, "wb") as f:
            pickle.dump({"precision": precision, "stats": stats}, f, pickle.HIGHEST_PROTOCOL)
        print("Wrote COCO eval results to: {}".format(eval_file))
        return stats

    def evaluate_detections(self, all_boxes, output_dir):
        os.makedirs(output_dir, exist_ok=True)
        res_file = os.path.join(output_dir, "detections_" + self._image_set + self._year + "_results")
        if self.config["use_salt"]:
            res_file += "_{}".format(str(uuid.uuid4()))
        res_file += ".json"
        self._write_coco_results_file(all_boxes, res_file)
        stats = None
        if "test" not in self._image_set:                      # test splits carry no ground truth
            stats = self._do_detection_eval(res_file, output_dir)
        if self.config["cleanup"]:
            os.remove(res_file)
        return stats

    def competition_mode(self, on):
        self.config["use_salt"] = not on
        self.config["cleanup"] = not on
