"""sys.path helper: puts this package's `lib/` (the mirror of the reference's lib/ import surface:
model, nets, layer_utils, nms, utils, datasets) and the optional tensorflow/matplotlib shims on sys.path,
the way the reference's tools/_init_paths.py:11-12 does for its own lib/."""
import os
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
SHIMS = os.path.join(PKG, "shims")


def add_lib_path(with_shims=False):
    for p in (ROOT, LIB):
        if p not in sys.path:
            sys.path.insert(0, p)
    if with_shims:
        import importlib.util
        for mod in ("tensorflow", "matplotlib"):
            if importlib.util.find_spec(mod) is None and SHIMS not in sys.path:
                sys.path.append(SHIMS)
