"""`py_cpu_nms(dets, thresh)`: the reference's readable baseline (lib/nms/py_cpu_nms.py:10-38) keeps a box while
`ovr <= thresh`, i.e. suppresses on `ovr > thresh` with '+1' areas in the dtype of `dets` -- for fp32 detections under
NumPy >= 2 that is exactly the gpu_nms predicate, which is what runs here (on the GPU; see nms/gpu_nms.py).
tests/test_oracle_golden.py pins this equivalence on outputs of the reference's own function."""
from nms.gpu_nms import gpu_nms


def py_cpu_nms(dets, thresh):
    return gpu_nms(dets, thresh)
