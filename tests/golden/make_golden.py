"""Generates tests/golden/reference_vectors.npz by IMPORTING the reference's own Python where it lies
(/root/reference/lib) -- the only reference modules importable without TensorFlow/easydict
(SURVEY.md 8(c)).  Run here (CPU container); the GPU box only reads the committed .npz.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/lib"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ga = load("layer_utils/generate_anchors.py", "ref_generate_anchors")
    pynms = load("nms/py_cpu_nms.py", "ref_py_cpu_nms")
    out = {}
    # (1) base anchors for the three anchor configurations of BASELINE.json
    for tag, scales in (("s3", (8, 16, 32)), ("s4", (4, 8, 16, 32)), ("s5", (2, 4, 8, 16, 32))):
        out["anchors_" + tag] = ga.generate_anchors(ratios=np.array((0.5, 1, 2)), scales=np.array(scales))
    # the MATLAB known-answer table in the comment of generate_anchors.py:14-39 (1-based => python + 1)
    out["anchors_matlab"] = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                                      [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                                      [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], np.float64)
    # (2) tiled anchors, numpy variant of snippets.py:14-30 restated inline is NOT reference code; instead pin the
    #     tiling through the formula's two ingredients that ARE reference outputs: base anchors (above).
    # (3) py_cpu_nms (the reference's readable baseline; predicate 'suppress when ovr > thresh'), tie-free scores
    rng = np.random.default_rng(20260922)
    for i, n in enumerate((1, 7, 64, 300, 1000)):
        xy = rng.uniform(0, 500, (n, 2)); wh = rng.uniform(10, 200, (n, 2))
        boxes = np.hstack([xy, xy + wh])
        boxes[n // 2:] = boxes[: n - n // 2] + rng.uniform(-8, 8, (n - n // 2, 4))
        scores = rng.permutation(n).astype(np.float64) / n + 0.001
        dets = np.hstack([boxes, scores[:, None]]).astype(np.float32)
        out["nms_dets_%d" % i] = dets
        for thr in (0.3, 0.7):
            out["nms_keep_%d_%d" % (i, int(thr * 10))] = np.asarray(pynms.py_cpu_nms(dets, thr), dtype=np.int64)
    # (4) im_list_to_blob (utils/blob.py:17-30) needs cv2 only at import time
    blob = load("utils/blob.py", "ref_blob")
    ims = [rng.standard_normal((37, 50, 3)).astype(np.float32), rng.standard_normal((30, 61, 3)).astype(np.float32)]
    out["blob_in0"], out["blob_in1"] = ims
    out["blob_out"] = blob.im_list_to_blob(ims)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; golden vectors are generated in the authoring container only")
    main()
