"""Pins the oracle (CPU restatement) against vectors produced by the reference's own importable Python
(tests/golden/make_golden.py) and cross-checks the C restatement against an independent numpy one."""
import os

import numpy as np
import pytest

from oracle import anchors as OA
from oracle import nms as ONMS
from oracle import boxes as OB

F = np.float32
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.mark.parametrize("tag,scales", [("s3", (8, 16, 32)), ("s4", (4, 8, 16, 32)), ("s5", (2, 4, 8, 16, 32))])
def test_base_anchors_match_reference(tag, scales):
    assert np.array_equal(OA.base_anchors(scales=scales), G["anchors_" + tag])


def test_base_anchors_known_answer_table():
    # lib/layer_utils/generate_anchors.py:14-39 (MATLAB, 1-based) == python + 1
    assert np.array_equal(OA.base_anchors() + 1, G["anchors_matlab"])


def test_tiled_anchor_order_and_values():
    a = OA.tiled_anchors(3, 5, 16, (8, 16, 32), (0.5, 1, 2))
    base = G["anchors_s3"]
    assert a.shape == (3 * 5 * 9, 4) and a.dtype == np.float32
    for (gy, gx, k) in [(0, 0, 0), (0, 1, 0), (1, 0, 8), (2, 4, 3)]:
        i = (gy * 5 + gx) * 9 + k
        assert np.array_equal(a[i], base[k] + np.array([gx * 16, gy * 16, gx * 16, gy * 16]))


@pytest.mark.parametrize("i", range(5))
@pytest.mark.parametrize("thr", [0.3, 0.7])
def test_plus1_strict_nms_matches_reference_py_cpu_nms(i, thr):
    dets = G["nms_dets_%d" % i]
    want = G["nms_keep_%d_%d" % (i, int(thr * 10))]
    assert np.array_equal(ONMS.nms_plus1_c(dets, thr, inclusive=False), want)
    assert np.array_equal(ONMS.nms_plus1_np(dets, thr, inclusive=False), want)


def _cases():
    rng = np.random.default_rng(5)
    for n in (1, 2, 63, 64, 65, 300, 1500):
        xy = rng.uniform(0, 300, (n, 2)); wh = rng.uniform(4, 120, (n, 2))
        b = np.round(np.hstack([xy, xy + wh])).astype(F)              # integer coords -> exact-threshold IoUs happen
        b[n // 2:] = b[: n - n // 2] + rng.integers(-3, 4, (n - n // 2, 4)).astype(F)
        s = (rng.integers(0, max(n // 3, 1), n) / max(n // 3, 1)).astype(F)   # many tied scores
        yield n, b, s


def test_c_restatement_equals_numpy_restatement():
    for n, b, s in _cases():
        d = np.hstack([b, s[:, None]]).astype(F)
        for thr in (0.3, 0.5, 0.7):
            for inc in (False, True):
                assert np.array_equal(ONMS.nms_plus1_c(d, thr, inc), ONMS.nms_plus1_np(d, thr, inc)), (n, thr, inc)
            for cap in (5, 300):
                assert np.array_equal(ONMS.nms_tf_c(b, s, cap, thr), ONMS.nms_tf_np(b, s, cap, thr)), (n, thr, cap)


def test_predicates_differ_exactly_at_threshold():
    d = np.array([[0, 0, 9, 9, 0.9], [0, 5, 9, 14, 0.8]], F)          # +1 IoU = 50/150
    t = float(F(50.0) / F(150.0))
    assert list(ONMS.nms_plus1_c(d, t, inclusive=False)) == [0, 1]
    assert list(ONMS.nms_plus1_c(d, t, inclusive=True)) == [0]


def test_threshold_double_to_float_rule():
    # cpu_nms compares an fp32 overlap with the DOUBLE 0.3: same as >= ceil32(0.3)
    t = ONMS.thresh_f32(0.3, inclusive=True)
    assert float(t) >= 0.3 and float(np.nextafter(t, F(0))) < 0.3
    assert ONMS.thresh_f32(0.5, True) == F(0.5)


def test_tf_nms_degenerate_and_cap():
    b = np.array([[0, 0, 10, 10], [5, 5, 5, 9], [0, 0, 10, 10], [20, 20, 10, 10], [11, 11, 19, 19]], F)
    s = np.array([0.9, 0.8, 0.7, 0.6, 0.5], F)
    assert list(ONMS.nms_tf_c(b, s, 10, 0.5)) == [0, 1, 3]             # inverted corners normalised (3 then suppresses 4); dup dropped
    assert list(ONMS.nms_tf_c(b, s, 2, 0.5)) == [0, 1]
    assert ONMS.nms_tf_c(np.zeros((0, 4), F), np.zeros(0, F), 5, 0.5).shape == (0,)


def test_tie_rule_lower_index_first():
    s = np.array([0.5, 0.7, 0.5, 0.7, 0.1], F)
    assert list(ONMS.argsort_desc(s)) == [1, 3, 0, 2, 4]


def test_box_codec_roundtrip_and_clips():
    rng = np.random.default_rng(0)
    b = np.array([[10, 20, 110, 70], [0, 0, 15, 15]], F)
    z = OB.decode(b, np.zeros((2, 8), F))
    # zero deltas: centre form of the +1 convention shifts x2,y2 by +1 (bbox_transform.py:41-63)
    assert np.allclose(z[:, :4], b + np.array([0, 0, 1, 1], F)) and np.array_equal(z[:, :4], z[:, 4:])
    d = OB.decode(b, rng.standard_normal((2, 8)).astype(F))
    c2 = OB.clip_two_sided(d, 60, 100)
    assert c2[:, 0::2].min() >= 0 and c2[:, 0::4].max() <= 99 and c2[:, 1::4].max() <= 59
    c1 = OB.clip_one_sided(d, 60, 100)
    assert (c1[:, 0::4] >= 0).all() and (c1[:, 2::4] <= 99).all()
    assert OB.decode(np.zeros((0, 4), F), np.zeros((0, 8), F)).shape == (0, 8)
