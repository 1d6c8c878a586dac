"""`nms(dets, thresh, force_cpu=False)` with the reference's contract (lib/model/nms_wrapper.py:15-23):
empty input -> []; otherwise indices into the UNSORTED `dets`, in descending-score order.

Both of the reference's predicates run on the GPU here (there is no CPU implementation):
  cfg.USE_GPU_NMS and not force_cpu -> gpu_nms semantics ('+1' areas, suppress when IoU >  thresh)
  otherwise                         -> cpu_nms semantics ('+1' areas, suppress when ovr >= thresh)
The host argsort mirrors gpu_nms.pyx:25-28 / cpu_nms.pyx:25 but is stable (ties: lower index first)."""
import numpy as np

from model.config import cfg
from tf_faster_rcnn_b200 import engine, ops


def nms(dets, thresh, force_cpu=False):
    if dets.shape[0] == 0:
        return []
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    order = np.argsort(-dets[:, 4], kind="stable")
    t32, flags = engine.nms_threshold(thresh, bool(cfg.USE_GPU_NMS) and not force_cpu)
    keep = ops.nms_host(dets[order], t32, flags, device_id=-1)   # the process's current device (rank-local under torchrun)
    return list(order[keep])
