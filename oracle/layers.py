"""Oracle: fp32 CPU restatement of the TF-1.x ops on the path (test infrastructure).

Activations are NHWC numpy fp32 (as in the reference graph); weights are HWIO as TF
stores them.  Convolutions and FC products run through torch-CPU fp32 (oneDNN) --
accumulation order is whatever that library does, exactly as the reference inherits
Eigen's.  TF op semantics restated here:
  conv SAME       pad_total = max((ceil(n/s)-1)*s + k - n, 0); before = pad_total//2
  conv2d_same     stride>1: explicit pad (k-1)//2, k-1-(k-1)//2 then VALID
                  (slim resnet_utils.conv2d_same; lib/nets/mobilenet_v1.py:21-49)
  max_pool SAME   padded cells never win
  batch_norm      tf.nn.batch_normalization inference: x*inv + (beta - mean*inv),
                  inv = gamma * rsqrt(var + eps)          (lib/nets/resnet_v1.py:22-44)
  softmax         exp(x - max) / sum
  crop_and_resize bilinear, extrapolation 0 (lib/nets/network.py:141-157)
"""
import numpy as np
import torch
import torch.nn.functional as TF

F = np.float32


def _t(x_nhwc):
    return torch.from_numpy(np.ascontiguousarray(x_nhwc)).permute(0, 3, 1, 2)


def _n(x_nchw):
    return np.ascontiguousarray(x_nchw.permute(0, 2, 3, 1).numpy())


def same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x, w_hwio, stride=1, padding="SAME", groups=1):
    """x NHWC fp32, w HWIO (or HW,C,1 depthwise via groups=C) -> NHWC fp32, no bias/act."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    xt = _t(x)
    if padding == "SAME":
        pt, pb = same_pads(x.shape[1], kh, stride)
        pl, pr = same_pads(x.shape[2], kw, stride)
        xt = TF.pad(xt, (pl, pr, pt, pb))
    if groups == 1:
        wt = torch.from_numpy(np.ascontiguousarray(w_hwio)).permute(3, 2, 0, 1).contiguous()
    else:  # depthwise: TF weight [kh,kw,C,1] -> torch [C,1,kh,kw]
        wt = torch.from_numpy(np.ascontiguousarray(w_hwio)).permute(2, 3, 0, 1).contiguous()
    return _n(TF.conv2d(xt, wt, None, stride=stride, padding=0, groups=groups))


def conv2d_same(x, w_hwio, stride, groups=1):
    """slim resnet_utils.conv2d_same / separable_conv2d_same."""
    if stride == 1:
        return conv2d(x, w_hwio, 1, "SAME", groups)
    k = w_hwio.shape[0]
    pb = (k - 1) // 2
    pe = k - 1 - pb
    xp = np.pad(x, ((0, 0), (pb, pe), (pb, pe), (0, 0)))
    return conv2d(xp, w_hwio, stride, "VALID", groups)


def bias_add(x, b):
    return x + b.astype(F)


def batch_norm(x, gamma, beta, mean, var, eps):
    inv = (gamma.astype(F) * (F(1.0) / np.sqrt(var.astype(F) + F(eps)))).astype(F)
    shift = (beta.astype(F) - mean.astype(F) * inv).astype(F)
    return x * inv + shift, inv, shift


def relu(x):
    return np.maximum(x, F(0))


def relu6(x):
    return np.minimum(np.maximum(x, F(0)), F(6))


def max_pool(x, k, s, padding):
    xt = _t(x)
    if padding == "SAME":
        pt, pb = same_pads(x.shape[1], k, s)
        pl, pr = same_pads(x.shape[2], k, s)
        xt = TF.pad(xt, (pl, pr, pt, pb), value=float("-inf"))
    return _n(TF.max_pool2d(xt, k, s))


def fully_connected(x, w_io):
    return (torch.from_numpy(np.ascontiguousarray(x)) @ torch.from_numpy(np.ascontiguousarray(w_io))).numpy()


def softmax_lastdim(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True, dtype=F)


def crop_and_resize(feat, boxes_yxyx, crop):
    """tf.image.crop_and_resize(feat[1,H,W,C], boxes[R,4] normalised y1,x1,y2,x2, box_ind=0,
    crop_size=[crop,crop]), bilinear, extrapolation_value 0.  All arithmetic fp32, op by op."""
    _, H, W, C = feat.shape
    R = boxes_yxyx.shape[0]
    b = boxes_yxyx.astype(F)
    y1, x1, y2, x2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    idx = np.arange(crop, dtype=F)
    hs = (y2 - y1) * F(H - 1) / F(crop - 1)
    ws = (x2 - x1) * F(W - 1) / F(crop - 1)
    in_y = (y1 * F(H - 1))[:, None] + idx[None, :] * hs[:, None]          # [R,crop]
    in_x = (x1 * F(W - 1))[:, None] + idx[None, :] * ws[:, None]
    vy = (in_y >= 0) & (in_y <= F(H - 1))
    vx = (in_x >= 0) & (in_x <= F(W - 1))
    iy = np.where(vy, in_y, F(0)); ix = np.where(vx, in_x, F(0))
    top = np.floor(iy).astype(np.int64); bot = np.ceil(iy).astype(np.int64)
    lef = np.floor(ix).astype(np.int64); rig = np.ceil(ix).astype(np.int64)
    yl = (iy - np.floor(iy)).astype(F)[:, :, None, None]
    xl = (ix - np.floor(ix)).astype(F)[:, None, :, None]
    f = feat[0]
    tl = f[top[:, :, None], lef[:, None, :]]
    tr = f[top[:, :, None], rig[:, None, :]]
    bl = f[bot[:, :, None], lef[:, None, :]]
    br = f[bot[:, :, None], rig[:, None, :]]
    t = tl + (tr - tl) * xl
    bo = bl + (br - bl) * xl
    out = t + (bo - t) * yl
    valid = (vy[:, :, None] & vx[:, None, :])[:, :, :, None]
    return np.where(valid, out, F(0)).astype(F)
