"""`get_imdb(name)` / `list_imdbs()` (lib/datasets/factory.py:17-52): the reference's names -- voc_{2007,2012}_{train,val,
trainval,test}[_diff], coco_2014_{train,val,minival,valminusminival,trainval}, coco_2015_{test,test-dev} -- resolve to
datasets.pascal_voc / datasets.coco (constructed lazily: they need the data on disk), plus image-directory and synthetic
imdbs that satisfy exactly what test_net needs (lib/model/test.py:138-192) when no dataset is available."""
import os
import pickle
import tempfile

import numpy as np

_SETS = {}


class SimpleImdb(object):
    def __init__(self, name, image_paths, num_classes):
        self._name = name
        self._paths = list(image_paths)
        self._classes = ["__background__"] + ["class_%d" % i for i in range(1, num_classes)]

    name = property(lambda self: self._name)
    num_classes = property(lambda self: len(self._classes))
    classes = property(lambda self: self._classes)
    image_index = property(lambda self: list(range(len(self._paths))))
    num_images = property(lambda self: len(self._paths))

    def image_path_at(self, i):
        return self._paths[i]

    def competition_mode(self, on):
        pass

    def image_id_at(self, i):
        return os.path.splitext(os.path.basename(self._paths[i]))[0]

    def write_voc_results(self, all_boxes, output_dir):
        """Per-class `comp4_det_test_<cls>.txt` in the VOCdevkit format (lib/datasets/pascal_voc.py:203-219):
        `<image id> <score %.3f> <x1+1 %.1f> <y1+1> <x2+1> <y2+1>` (the devkit is 1-based)."""
        files = []
        for j, cls in enumerate(self._classes):
            if j == 0:
                continue
            path = os.path.join(output_dir, "comp4_det_test_%s.txt" % cls)
            with open(path, "wt") as f:
                for i in range(len(self._paths)):
                    dets = all_boxes[j][i]
                    for k in range(len(dets)):
                        f.write("{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n".format(
                            self.image_id_at(i), dets[k][-1], dets[k][0] + 1, dets[k][1] + 1, dets[k][2] + 1, dets[k][3] + 1))
            files.append(path)
        return files

    def write_coco_results(self, all_boxes, res_file):
        """COCO results json (lib/datasets/coco.py:258-292): bbox = [x, y, w, h] with w = x2 - x1 + 1, h = y2 - y1 + 1."""
        import json
        results = []
        for j in range(1, len(self._classes)):
            for i in range(len(self._paths)):
                dets = np.asarray(all_boxes[j][i], dtype=np.float64)
                for k in range(len(dets)):
                    x1, y1, x2, y2, sc = dets[k][:5]
                    results.append({"image_id": self.image_id_at(i), "category_id": j,
                                    "bbox": [float(x1), float(y1), float(x2 - x1 + 1), float(y2 - y1 + 1)], "score": float(sc)})
        with open(res_file, "w") as f:
            json.dump(results, f)
        return results

    def evaluate_detections(self, all_boxes, output_dir=None):
        """No ground truth exists for these sets: writes the VOC-format and COCO-format result files plus a per-class
        detection count summary instead of computing AP."""
        counts = [int(sum(len(d) for d in per_image)) for per_image in all_boxes]
        if output_dir:
            with open(os.path.join(output_dir, "detection_counts.pkl"), "wb") as f:
                pickle.dump(counts, f)
            self.write_voc_results(all_boxes, output_dir)
            self.write_coco_results(all_boxes, os.path.join(output_dir, "detections_%s_results.json" % self._name))
        print("detections per class:", counts[1:])
        return counts


def _synthetic(n, num_classes, h=375, w=500, seed=3):
    import cv2
    d = tempfile.mkdtemp(prefix="frcnn_synth_")
    rng = np.random.default_rng(seed)
    paths = []
    for i in range(n):
        im = cv2.blur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (5, 5))
        p = os.path.join(d, "%06d.png" % i)
        cv2.imwrite(p, im)
        paths.append(p)
    return SimpleImdb("synthetic_%d_%d" % (n, num_classes), paths, num_classes)


def _voc(split, year, use_diff=False):
    from datasets.pascal_voc import pascal_voc
    return pascal_voc(split, year, use_diff=use_diff)


def _coco(split, year):
    from datasets.coco import coco
    return coco(split, year)


for _year in ("2007", "2012"):
    for _split in ("train", "val", "trainval", "test"):
        _SETS["voc_{}_{}".format(_year, _split)] = (lambda split=_split, year=_year: _voc(split, year))
        _SETS["voc_{}_{}_diff".format(_year, _split)] = (lambda split=_split, year=_year: _voc(split, year, True))
for _split in ("train", "val", "minival", "valminusminival", "trainval"):
    _SETS["coco_2014_{}".format(_split)] = (lambda split=_split: _coco(split, "2014"))
for _split in ("test", "test-dev"):
    _SETS["coco_2015_{}".format(_split)] = (lambda split=_split: _coco(split, "2015"))


def get_imdb(name):
    """A registered name (reference sets above, register()) | 'synthetic_<n>_<classes>' | 'dir:<path>:<classes>'."""
    if name in _SETS:
        return _SETS[name]()
    if name.startswith("synthetic_"):
        _, n, c = name.split("_")
        return _synthetic(int(n), int(c))
    if name.startswith("dir:"):
        _, path, c = name.split(":")
        files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.lower().endswith((".jpg", ".jpeg", ".png")))
        return SimpleImdb("dir_" + os.path.basename(os.path.normpath(path)), files, int(c))
    raise KeyError('Unknown dataset: {}'.format(name))


def register(name, fn):
    _SETS[name] = fn


def list_imdbs():
    return list(_SETS.keys()) + ["synthetic_<n>_<classes>", "dir:<path>:<classes>"]
