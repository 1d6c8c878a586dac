"""`gpu_nms(dets, thresh, device_id=0)` -- the Cython entry of the reference (lib/nms/gpu_nms.pyx:16-31) on the B200 path:
score argsort on the host, the suppression in tf_faster_rcnn_b200/csrc/nms.cu through the `_nms`-compatible C entry
(`frcnn_nms_host`), predicate of nms_kernel.cu ('+1' areas, suppress when IoU > float(thresh)).
Ties in score keep the lower index first (stable sort); the reference's `argsort()[::-1]` leaves that order unspecified."""
import numpy as np

from tf_faster_rcnn_b200 import engine, ops


def gpu_nms(dets, thresh, device_id=0):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    order = np.argsort(-dets[:, 4], kind="stable")
    t32, flags = engine.nms_threshold(thresh, True)
    keep = ops.nms_host(dets[order], t32, flags, device_id=device_id)
    return list(order[keep])
