"""Image-database base class: the part of the reference's `imdb` (lib/datasets/imdb.py:20-125) the TEST path and its
consumers use -- identity, class list, image index -> path, ground-truth roidb, `evaluate_detections`,
`competition_mode`.  Training-only helpers (flipping, proposal-recall evaluation, roidb merging) are out of scope."""
import os

from model.config import cfg


class imdb(object):
    def __init__(self, name, classes=None):
        self._name = name
        self._classes = tuple(classes) if classes else ()
        self._image_index = []
        self._roidb = None
        self.config = {}

    name = property(lambda self: self._name)
    classes = property(lambda self: self._classes)
    num_classes = property(lambda self: len(self._classes))
    image_index = property(lambda self: self._image_index)
    num_images = property(lambda self: len(self._image_index))

    @property
    def cache_path(self):
        """<DATA_DIR>/cache, created on first use (imdb.py:85-90)."""
        path = os.path.abspath(os.path.join(cfg.DATA_DIR, "cache"))
        os.makedirs(path, exist_ok=True)
        return path

    @property
    def roidb(self):
        """Per-image ground truth: list of {boxes uint16 [n,4] 0-based, gt_classes int32 [n], gt_overlaps csr [n,C],
        flipped False, seg_areas float32 [n]}; built once by `gt_roidb()` (imdb.py:62-71)."""
        if self._roidb is None:
            self._roidb = self.gt_roidb()
        return self._roidb

    def image_path_at(self, i):
        raise NotImplementedError

    def image_id_at(self, i):
        return self._image_index[i]

    def gt_roidb(self):
        raise NotImplementedError

    def evaluate_detections(self, all_boxes, output_dir=None):
        """all_boxes[class][image] = float array [n, 5] (x1, y1, x2, y2, score), as test_net builds it
        (lib/model/test.py:145-181)."""
        raise NotImplementedError

    def competition_mode(self, on):
        pass
