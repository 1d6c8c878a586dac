"""`cpu_nms(dets, thresh)` with the predicate of the reference's Cython routine (lib/nms/cpu_nms.pyx:17-68: '+1' areas,
fp32 overlap compared against the DOUBLE threshold with >=).  The name is the reference's; the work runs on the GPU
(tf_faster_rcnn_b200/csrc/nms.cu) -- this build has no CPU implementation of any stage."""
import numpy as np

from tf_faster_rcnn_b200 import engine, ops


def cpu_nms(dets, thresh):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    order = np.argsort(-dets[:, 4], kind="stable")
    t32, flags = engine.nms_threshold(thresh, False)
    keep = ops.nms_host(dets[order], t32, flags, device_id=-1)
    return list(order[keep])
