"""Oracle NMS front-end (test infrastructure, see oracle/__init__.py).

`*_c` functions call the plain-C restatement (oracle/nms_c.c, built by oracle/build.py);
`*_np` functions are independent vectorised-numpy restatements used to cross-check the C
code on small inputs.  Semantics table: SURVEY.md section 8(a) row N.
"""
import ctypes
import os
import numpy as np

F = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_nms.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build_c()
        L = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        L.oracle_nms_plus1.argtypes = [fp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ip]
        L.oracle_nms_plus1.restype = ctypes.c_int
        L.oracle_nms_tf.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ip]
        L.oracle_nms_tf.restype = ctypes.c_int
        L.oracle_argsort_desc.argtypes = [fp, ctypes.c_int, ip]
        L.oracle_argsort_desc.restype = None
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def thresh_f32(thresh, inclusive):
    """The reference compares an fp32 overlap with a *double* threshold in cpu_nms
    (cpu_nms.pyx:17 `np.float thresh`, :65 `ovr >= thresh`) and with a float in
    nms_kernel.cu:34,71.  For fp32 ovr, `ovr >= t64` <=> `ovr >= ceil32(t64)` and
    `ovr > t64` <=> `ovr > floor32(t64)`; the CUDA kernel uses float(t64) = RN."""
    t32 = F(thresh)
    if inclusive:
        if float(t32) < float(thresh):
            t32 = np.nextafter(t32, F(np.inf))
    return F(t32)


def argsort_desc(scores):
    scores = np.ascontiguousarray(scores, dtype=F).ravel()
    out = np.empty(scores.shape[0], dtype=np.int32)
    lib().oracle_argsort_desc(_fp(scores), scores.shape[0], _ip(out))
    return out


def nms_plus1_c(dets, thresh, inclusive):
    """cpu_nms (inclusive=True) / gpu_nms & py_cpu_nms predicate (inclusive=False)."""
    dets = np.ascontiguousarray(dets, dtype=F)
    n = dets.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    keep = np.empty(n, dtype=np.int32)
    k = lib().oracle_nms_plus1(_fp(dets), n, float(thresh_f32(thresh, inclusive)), int(bool(inclusive)), _ip(keep))
    return keep[:k].copy()


def nms_tf_c(boxes, scores, max_out, thr):
    boxes = np.ascontiguousarray(boxes, dtype=F)
    scores = np.ascontiguousarray(scores, dtype=F).ravel()
    n = boxes.shape[0]
    keep = np.empty(max(max_out, 1), dtype=np.int32)
    k = lib().oracle_nms_tf(_fp(boxes), _fp(scores), n, int(max_out), float(F(thr)), _ip(keep))
    return keep[:k].copy()


# ---- independent numpy restatements (small inputs only) ---------------------------------

def _order(scores):
    scores = np.asarray(scores, dtype=F)
    return np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64))).astype(np.int32)


def nms_plus1_np(dets, thresh, inclusive):
    dets = np.asarray(dets, dtype=F)
    n = dets.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    t = thresh_f32(thresh, inclusive)
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    area = (x2 - x1 + F(1)) * (y2 - y1 + F(1))
    order = _order(dets[:, 4])
    dead = np.zeros(n, dtype=bool)
    keep = []
    for a in range(n):
        i = order[a]
        if dead[i]:
            continue
        keep.append(i)
        rest = order[a + 1:]
        w = np.maximum(F(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + F(1))
        h = np.maximum(F(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + F(1))
        inter = w * h
        ovr = inter / (area[i] + area[rest] - inter)
        dead[rest[(ovr >= t) if inclusive else (ovr > t)]] = True
    return np.asarray(keep, dtype=np.int32)


def nms_tf_np(boxes, scores, max_out, thr):
    boxes = np.asarray(boxes, dtype=F)
    thr = F(thr)
    order = _order(scores)
    y0 = np.minimum(boxes[:, 0], boxes[:, 2]); x0 = np.minimum(boxes[:, 1], boxes[:, 3])
    y1 = np.maximum(boxes[:, 0], boxes[:, 2]); x1 = np.maximum(boxes[:, 1], boxes[:, 3])
    area = (y1 - y0) * (x1 - x0)
    keep = []
    for i in order:
        if len(keep) >= max_out:
            break
        if keep:
            k = np.asarray(keep)
            inter = (np.maximum(np.minimum(y1[i], y1[k]) - np.maximum(y0[i], y0[k]), F(0)) *
                     np.maximum(np.minimum(x1[i], x1[k]) - np.maximum(x0[i], x0[k]), F(0)))
            with np.errstate(divide="ignore", invalid="ignore"):
                iou = inter / (area[i] + area[k] - inter)
            iou = np.where((area[i] <= 0) | (area[k] <= 0), F(0), iou)
            if np.any(iou > thr):
                continue
        keep.append(int(i))
    return np.asarray(keep, dtype=np.int32)
