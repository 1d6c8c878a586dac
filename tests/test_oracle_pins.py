"""Independent pins for the TensorFlow-side ops of the oracle (VERDICT r01 weak #1 / next #2d).

TensorFlow 1.x is not installable offline, so the oracle's conv / BN / pool / softmax / crop_and_resize /
non_max_suppression / top_k cannot be compared with TF itself.  Each of them is instead compared with an implementation
that does NOT share code (or author) with oracle/: scipy.signal / scipy.ndimage, torchvision.ops.nms,
torch.nn.functional.grid_sample, torch float64 reductions, and scalar loops transcribed from the published TF kernel
formulas (tensorflow/core/kernels/crop_and_resize_op.cc, non_max_suppression_op.cc, the "SAME" padding rule of the
tf.nn convolution notes).  CPU only, seconds.
"""
import numpy as np
import pytest
import scipy.ndimage
import scipy.signal
import torch

from oracle import layers as L
from oracle import nms as ONMS
from oracle import pipeline as P

F = np.float32


# ---------------------------------------------------------------------------------------------------------------------
# conv2d: scipy.signal.correlate ('same' centring, stride 1) and a scalar-formula transcription for strided SAME/VALID
def _conv_scalar_f64(x, w, stride, pads):
    """out[n,i,j,co] = sum x[n, i*s + r - pt, j*s + q - pl, ci] * w[r,q,ci,co]: float64 loops over taps only."""
    (pt, pb), (pl, pr) = pads
    xp = np.pad(x.astype(np.float64), ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    kh, kw = w.shape[:2]
    ho = (xp.shape[1] - kh) // stride + 1
    wo = (xp.shape[2] - kw) // stride + 1
    out = np.zeros((x.shape[0], ho, wo, w.shape[3]))
    for r in range(kh):
        for q in range(kw):
            patch = xp[:, r:r + (ho - 1) * stride + 1:stride, q:q + (wo - 1) * stride + 1:stride, :]
            out += np.einsum("nhwc,co->nhwo", patch, w[r, q].astype(np.float64))
    return out


def _tf_same(n, k, s):
    """tf.nn 'Notes on padding': out = ceil(n/s); pad_along = max((out-1)*s + k - n, 0); pad_before = pad_along // 2."""
    out = (n + s - 1) // s
    pad = max((out - 1) * s + k - n, 0)
    return pad // 2, pad - pad // 2


@pytest.mark.parametrize("k", [1, 3, 7])
def test_conv_same_stride1_matches_scipy_correlate(k):
    rng = np.random.default_rng(k)
    x = rng.standard_normal((1, 11, 13, 3)).astype(F)
    w = rng.standard_normal((k, k, 3, 2)).astype(F)
    got = L.conv2d(x, w, 1, "SAME")
    want = np.zeros((11, 13, 2))
    for co in range(2):
        for ci in range(3):
            want[:, :, co] += scipy.signal.correlate2d(x[0, :, :, ci].astype(np.float64), w[:, :, ci, co].astype(np.float64), mode="same")
    assert np.abs(got[0] - want).max() < 2e-5 * np.abs(want).max()


@pytest.mark.parametrize("n,k,s", [(4, 3, 2), (5, 3, 2), (7, 7, 2), (600, 7, 2), (38, 3, 1), (75, 3, 2), (10, 1, 2)])
def test_same_padding_rule_known_answers(n, k, s):
    assert L.same_pads(n, k, s) == _tf_same(n, k, s)


def test_same_padding_tiny_known_answer():
    # n=4, k=3, s=2, unit weights: TF pads (0, 1): out = [x0+x1+x2, x2+x3]
    x = np.array([1, 10, 100, 1000], F).reshape(1, 1, 4, 1)
    w = np.ones((1, 3, 1, 1), F)
    assert L.conv2d(x, w, 2, "SAME").ravel().tolist() == [111.0, 1100.0]
    # slim conv2d_same (explicit pad (k-1)//2 = 1 on both sides, then VALID): out = [0+x0+x1, x1+x2+x3]
    wc = np.zeros((3, 3, 1, 1), F); wc[:, 1] = 1              # only the centre column is non-zero: a 1-D filter along H
    assert L.conv2d_same(x.reshape(1, 4, 1, 1), wc, 2)[0, :, 0, 0].tolist() == [11.0, 1110.0]


@pytest.mark.parametrize("stride,mode", [(2, "SAME"), (2, "EXPLICIT"), (1, "SAME")])
def test_conv_strided_matches_scalar_formula(stride, mode):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 9, 12, 5)).astype(F)
    w = rng.standard_normal((3, 3, 5, 4)).astype(F)
    if mode == "SAME":
        got = L.conv2d(x, w, stride, "SAME")
        pads = (_tf_same(9, 3, stride), _tf_same(12, 3, stride))
    else:
        got = L.conv2d_same(x, w, stride)
        pads = ((1, 1), (1, 1))
    want = _conv_scalar_f64(x, w, stride, pads)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5 * np.abs(want).max()


def test_depthwise_matches_per_channel_correlate():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 8, 9, 4)).astype(F)
    w = rng.standard_normal((3, 3, 4, 1)).astype(F)
    got = L.conv2d(x, w, 1, "SAME", groups=4)
    for c in range(4):
        want = scipy.signal.correlate2d(x[0, :, :, c].astype(np.float64), w[:, :, c, 0].astype(np.float64), mode="same")
        assert np.abs(got[0, :, :, c] - want).max() < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
def test_batch_norm_matches_textbook_float64():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 5, 6, 8)).astype(F) * 3
    g, b = rng.uniform(0.5, 2, 8).astype(F), rng.standard_normal(8).astype(F)
    m, v = rng.standard_normal(8).astype(F), rng.uniform(0.1, 4, 8).astype(F)
    got, inv, shift = L.batch_norm(x, g, b, m, v, 1e-5)
    want = g.astype(np.float64) * (x.astype(np.float64) - m) / np.sqrt(v.astype(np.float64) + 1e-5) + b
    assert np.abs(got - want).max() < 1e-5


def test_softmax_matches_torch_float64():
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((300, 81)) * 8).astype(F)
    want = torch.softmax(torch.from_numpy(x.astype(np.float64)), -1).numpy()
    got = L.softmax_lastdim(x)
    assert np.abs(got - want).max() < 5e-7 and got.dtype == np.float32        # fp32 exp + one division
    # the RPN's 2-way softmax (network.py:68-86)
    two = rng.standard_normal((1, 4, 5, 18)).astype(F)
    fg = P.rpn_fg_prob(two)
    # TF graph: reshape [1,h,w,2A] -> [1, 2, A*h... ] pairs (bg_a, fg_a) = channels (a, A + a); softmax over the pair
    z = torch.from_numpy(two.astype(np.float64))
    want_fg = torch.softmax(torch.stack([z[..., :9], z[..., 9:]], -1), -1)[..., 1].numpy()
    assert fg.size == two.size // 2 and np.abs(fg.reshape(want_fg.shape) - want_fg).max() < 5e-7


@pytest.mark.parametrize("k,s,mode", [(2, 2, "SAME"), (3, 2, "SAME"), (2, 2, "VALID"), (3, 2, "VALID")])
def test_max_pool_matches_brute_force(k, s, mode):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((1, 9, 11, 3)).astype(F)
    got = L.max_pool(x, k, s, mode)
    if mode == "SAME":
        (pt, pb), (pl, pr) = _tf_same(9, k, s), _tf_same(11, k, s)
    else:
        pt = pb = pl = pr = 0
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)), constant_values=-np.inf)
    ho, wo = (xp.shape[1] - k) // s + 1, (xp.shape[2] - k) // s + 1
    want = np.empty((1, ho, wo, 3), F)
    for i in range(ho):
        for j in range(wo):
            want[0, i, j] = xp[0, i * s:i * s + k, j * s:j * s + k].reshape(-1, 3).max(0)
    assert np.array_equal(got, want)


def test_max_pool_stride1_matches_scipy_maximum_filter():
    x = np.random.default_rng(17).standard_normal((1, 8, 10, 2)).astype(F)
    got = L.max_pool(x, 3, 1, "SAME")
    want = scipy.ndimage.maximum_filter(x[0], size=(3, 3, 1), mode="constant", cval=-np.inf)
    assert np.array_equal(got[0], want)


# ---------------------------------------------------------------------------------------------------------------------
def _crop_and_resize_scalar(feat, boxes, crop):
    """crop_and_resize_op.cc (CPU functor), transcribed loop for loop, float32 scalars."""
    _, H, W, C = feat.shape
    out = np.zeros((boxes.shape[0], crop, crop, C), F)
    for b in range(boxes.shape[0]):
        y1, x1, y2, x2 = (F(v) for v in boxes[b])
        height_scale = (y2 - y1) * F(H - 1) / F(crop - 1) if crop > 1 else F(0)
        width_scale = (x2 - x1) * F(W - 1) / F(crop - 1) if crop > 1 else F(0)
        for y in range(crop):
            in_y = y1 * F(H - 1) + F(y) * height_scale if crop > 1 else F(0.5) * (y1 + y2) * F(H - 1)
            if in_y < 0 or in_y > H - 1:
                continue                                      # extrapolation_value = 0
            top, bottom = int(np.floor(in_y)), int(np.ceil(in_y))
            y_lerp = F(in_y - F(top))
            for x in range(crop):
                in_x = x1 * F(W - 1) + F(x) * width_scale if crop > 1 else F(0.5) * (x1 + x2) * F(W - 1)
                if in_x < 0 or in_x > W - 1:
                    continue
                left, right = int(np.floor(in_x)), int(np.ceil(in_x))
                x_lerp = F(in_x - F(left))
                tl, tr = feat[0, top, left], feat[0, top, right]
                bl, br = feat[0, bottom, left], feat[0, bottom, right]
                t = tl + (tr - tl) * x_lerp
                bo = bl + (br - bl) * x_lerp
                out[b, y, x] = t + (bo - t) * y_lerp
    return out


def test_crop_and_resize_matches_tf_kernel_transcription():
    rng = np.random.default_rng(8)
    feat = rng.standard_normal((1, 12, 15, 6)).astype(F)
    boxes = np.array([[0.1, 0.2, 0.7, 0.9], [0.0, 0.0, 1.0, 1.0], [-0.2, 0.3, 0.5, 1.3], [0.4, 0.4, 0.4, 0.4], [0.9, 0.1, 0.2, 0.8],
                      [0.33, 0.5, 1.0, 0.75]], F)
    for crop in (7, 14):
        assert np.array_equal(L.crop_and_resize(feat, boxes, crop), _crop_and_resize_scalar(feat, boxes, crop))


def test_crop_and_resize_matches_torch_grid_sample_inside_the_map():
    """for boxes inside the map crop_and_resize == bilinear sampling at in = b1*(n-1) + i*(b2-b1)*(n-1)/(crop-1): that is
    grid_sample(align_corners=True) on the same points (an implementation that shares nothing with TF or the oracle)."""
    rng = np.random.default_rng(9)
    H, W, C, crop = 10, 14, 4, 7
    feat = rng.standard_normal((1, H, W, C)).astype(F)
    boxes = rng.uniform(0.05, 0.95, (16, 4)).astype(F)
    boxes = np.stack([np.minimum(boxes[:, 0], boxes[:, 2]), np.minimum(boxes[:, 1], boxes[:, 3]),
                      np.maximum(boxes[:, 0], boxes[:, 2]), np.maximum(boxes[:, 1], boxes[:, 3])], 1)
    got = L.crop_and_resize(feat, boxes, crop)
    t = np.linspace(0, 1, crop)
    ys = boxes[:, 0:1] + t[None, :] * (boxes[:, 2:3] - boxes[:, 0:1])          # normalised [0,1] -> grid [-1,1]
    xs = boxes[:, 1:2] + t[None, :] * (boxes[:, 3:4] - boxes[:, 1:2])
    grid = np.stack(np.broadcast_arrays(2 * xs[:, None, :] - 1, 2 * ys[:, :, None] - 1), -1)   # [R,crop,crop,(x,y)]
    ft = torch.from_numpy(feat.astype(np.float64)).permute(0, 3, 1, 2).expand(16, -1, -1, -1)
    want = torch.nn.functional.grid_sample(ft, torch.from_numpy(grid), mode="bilinear", align_corners=True).permute(0, 2, 3, 1).numpy()
    assert np.abs(got - want).max() < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
def _random_boxes(rng, n):
    xy = rng.uniform(0, 500, (n, 2))
    wh = rng.uniform(8, 200, (n, 2))
    return np.hstack([xy, xy + wh]).astype(F)


@pytest.mark.parametrize("n,thr", [(50, 0.7), (600, 0.7), (600, 0.3), (3000, 0.5)])
def test_tf_nms_matches_torchvision_on_regular_boxes(n, thr):
    """tf.image.non_max_suppression (strict '>', continuous areas, score order, cap) vs torchvision.ops.nms -- an
    implementation from a different code base with the same published predicate (IoU > thr suppresses)."""
    import torchvision
    rng = np.random.default_rng(n)
    b = _random_boxes(rng, n)
    s = rng.permutation(n).astype(F) / n                   # distinct scores: order is unambiguous
    want = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
    for cap in (300, n):
        got = ONMS.nms_tf_c(b, s, cap, thr)
        assert np.array_equal(got, want[:cap]), (n, thr, cap)


def test_tf_nms_matches_kernel_transcription_with_ties_and_degenerate_boxes():
    """non_max_suppression_op.cc: candidates by descending score (ties: lower index), IoU against the SELECTED set only,
    IoU := 0 when either area <= 0, suppress when IoU > thr; coordinates may be given in either corner order."""
    rng = np.random.default_rng(4)
    b = np.round(_random_boxes(rng, 400) / 8).astype(F) * 8
    b[::7, 2] = b[::7, 0]                                   # zero-width boxes
    b[5::11] = b[5::11][:, [2, 3, 0, 1]]                    # flipped corners
    s = (rng.integers(0, 40, 400) / 40).astype(F)

    def iou(a, c):
        ay0, ax0, ay1, ax1 = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
        cy0, cx0, cy1, cx1 = min(c[0], c[2]), min(c[1], c[3]), max(c[0], c[2]), max(c[1], c[3])
        aa, ac = F(ay1 - ay0) * F(ax1 - ax0), F(cy1 - cy0) * F(cx1 - cx0)
        if aa <= 0 or ac <= 0:
            return F(0)
        ih = max(F(min(ay1, cy1) - max(ay0, cy0)), F(0)); iw = max(F(min(ax1, cx1) - max(ax0, cx0)), F(0))
        inter = F(ih * iw)
        return F(inter / F(F(aa + ac) - inter))
    order = sorted(range(400), key=lambda i: (-float(s[i]), i))
    keep = []
    for i in order:
        if len(keep) == 120:
            break
        if all(not (iou(b[i], b[j]) > F(0.5)) for j in keep):
            keep.append(i)
    assert list(ONMS.nms_tf_c(b, s, 120, 0.5)) == keep
    assert list(ONMS.nms_tf_np(b, s, 120, 0.5)) == keep


def test_top_k_order_matches_torch_stable_sort():
    rng = np.random.default_rng(12)
    s = (rng.integers(0, 50, 5000) / 50).astype(F)         # heavy ties: lower index first (tf.nn.top_k contract)
    want = torch.sort(torch.from_numpy(s), descending=True, stable=True).indices.numpy()
    assert np.array_equal(ONMS.argsort_desc(s), want)
    o = P.opts(test_mode="top", rpn_top_n=300)
    props = rng.standard_normal((5000, 4)).astype(F)
    rois, sc, keep = P.proposals_top(s, props, o)
    assert np.array_equal(keep, want[:300]) and np.array_equal(rois[:, 1:], props[want[:300]])


def test_plus1_nms_matches_torchvision_after_the_plus1_shift():
    """'+1' pixel-area IoU of (x1,y1,x2,y2) == continuous IoU of (x1,y1,x2+1,y2+1): torchvision on shifted boxes is an
    independent check of cpu_nms / gpu_nms (strict variant) away from exact-threshold ties."""
    import torchvision
    rng = np.random.default_rng(21)
    b = np.round(_random_boxes(rng, 800)).astype(F)
    s = rng.permutation(800).astype(F) / 800
    shifted = b.copy(); shifted[:, 2:] += 1
    for thr in (0.3, 0.5, 0.7):
        want = torchvision.ops.nms(torch.from_numpy(shifted), torch.from_numpy(s), thr).numpy()
        d = np.hstack([b, s[:, None]]).astype(F)
        got = ONMS.nms_plus1_c(d, thr, inclusive=False)
        # torchvision divides in a different order; only a tie within 1 ulp of thr could differ, none in this seed
        assert np.array_equal(np.sort(got), np.sort(want)), thr
