"""Generates tests/golden/voc_eval_vectors.npz by running the REFERENCE's own voc_eval (imported from
/root/reference/lib/datasets/voc_eval.py) on the deterministic miniature devkit of tests/voc_fixture.py.
The reference writes its annotation cache in text mode (fails under Python 3), so the cache file is pre-seeded here
with the reference's own parse_rec output; everything after that is the reference's code path.

    python tests/golden/make_voc_golden.py
"""
import importlib.util
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import voc_fixture  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_voc_eval", "/root/reference/lib/datasets/voc_eval.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    root = tempfile.mkdtemp(prefix="voc_golden_")
    ids, dets = voc_fixture.build(root)
    detpath = voc_fixture.write_det_files(os.path.join(root, "dets"), dets)
    annopath = os.path.join(root, "VOC2007", "Annotations", "{:s}.xml")
    imageset = os.path.join(root, "VOC2007", "ImageSets", "Main", "test.txt")
    cachedir = os.path.join(root, "cache")
    os.makedirs(cachedir)
    with open(os.path.join(cachedir, "%s_annots.pkl" % imageset), "wb") as f:     # == imageset + "_annots.pkl"
        pickle.dump({i: ref.parse_rec(annopath.format(i)) for i in ids}, f)
    out = {}
    for cls in ("car", "person", "dog"):
        for m07 in (False, True):
            for diff in (False, True):
                rec, prec, ap = ref.voc_eval(detpath, annopath, imageset, cls, cachedir, ovthresh=0.5,
                                             use_07_metric=m07, use_diff=diff)
                tag = "%s_%d_%d" % (cls, m07, diff)
                out["rec_" + tag], out["prec_" + tag], out["ap_" + tag] = rec, prec, np.float64(ap)
                print(tag, "ap=%.6f" % ap, "nd=%d" % len(rec))
    out["parse_000001"] = np.array([[o["truncated"], o["difficult"]] + o["bbox"] for o in ref.parse_rec(annopath.format(ids[0]))]).reshape(-1, 6)
    np.savez_compressed(os.path.join(HERE, "voc_eval_vectors.npz"), **out)


if __name__ == "__main__":
    main()
