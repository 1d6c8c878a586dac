"""Put this repo's lib/ mirror (+ tensorflow/matplotlib shims when the real packages are absent) on sys.path."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tf_faster_rcnn_b200 import paths  # noqa: E402

paths.add_lib_path(with_shims=True)
