"""Parity of the tcgen05 implicit-GEMM conv/FC kernel (FP16x3; the r01 3xTF32 kernel via impl=1) against the oracle's fp32 CPU conv and a
float64 reference.  Tolerance: the GPU result must be as close to the float64 truth as fp32 arithmetic allows
(<= 4e-6 of the output's max magnitude; the fp32 oracle itself sits at ~1e-6) -- written per test."""
import numpy as np
import pytest
import torch

from oracle import layers as L

pytestmark = pytest.mark.gpu
F = np.float32


def ref64(x, w, stride, pad_t, pad_l, ho, wo):
    xt = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w.astype(np.float64)).permute(3, 2, 0, 1)
    kh, kw = w.shape[:2]
    h, wd = x.shape[1:3]
    pb = max((ho - 1) * stride + kh - h - pad_t, 0)
    pr = max((wo - 1) * stride + kw - wd - pad_l, 0)
    xt = torch.nn.functional.pad(xt, (pad_l, pr, pad_t, pb))
    y = torch.nn.functional.conv2d(xt, wt, None, stride=stride)
    return y.permute(0, 2, 3, 1).numpy()[:, :ho, :wo]


def run_conv(x, w, stride, mode, scale=None, shift=None, residual=None, act=0, block_n=0, kb_per_chunk=0, time_it=False, split_k=0):
    from tf_faster_rcnn_b200 import ops
    n, h, wd, cin = x.shape
    k = w.shape[0]
    ho, wo, pt, pl = ops.conv_out_hw(h, wd, k, stride, mode)
    pc = ops.PackedConv(w, scale, shift)
    xd = torch.from_numpy(x).cuda()
    out = torch.full((n, ho, wo, w.shape[3]), float("nan"), dtype=torch.float32, device="cuda")
    rd = None if residual is None else torch.from_numpy(residual).cuda()
    plan = ops.ConvPlan(xd, pc, out, stride, pt, pl, act, rd, block_n, kb_per_chunk, split_k)
    plan.run()
    torch.cuda.synchronize()
    info = plan.info()
    if time_it:
        for _ in range(3):
            plan.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        info["us"] = e0.elapsed_time(e1) * 1000 / 20
        info["tflops"] = 2.0 * n * ho * wo * w.shape[3] * k * k * cin / (info["us"] * 1e-6) / 1e12
    return out.cpu().numpy(), info, (ho, wo, pt, pl)


CASES = [
    # name, n, h, w, cin, cout, k, stride, mode, block_n
    ("fc_small", 1, 1, 300, 64, 64, 1, 1, "SAME", 0),
    ("pw_1900_bn64", 1, 38, 50, 256, 256, 1, 1, "SAME", 64),
    ("pw_1900_bn128", 1, 38, 50, 1024, 256, 1, 1, "SAME", 128),
    ("c3_s1_64", 1, 38, 50, 64, 64, 3, 1, "SAME", 0),
    ("c3_s1_odd_cout", 1, 20, 30, 128, 96, 3, 1, "SAME", 0),
    ("c3_s1_big", 1, 75, 100, 64, 128, 3, 1, "SAME", 0),
    ("c3_s2_explicit", 1, 75, 100, 64, 64, 3, 2, "EXPLICIT", 0),
    ("c3_s2_odd", 1, 38, 51, 32, 64, 3, 2, "EXPLICIT", 0),
    ("head_c3_rois", 20, 7, 7, 64, 64, 3, 1, "SAME", 0),
    ("head_pw_rois", 20, 7, 7, 128, 64, 1, 1, "SAME", 0),
    ("fc_k3136", 1, 1, 300, 3136, 128, 1, 1, "SAME", 0),
    ("fc_cout405", 1, 1, 300, 256, 405, 1, 1, "SAME", 0),
    ("rpn_cout72", 1, 38, 50, 512, 72, 1, 1, "SAME", 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_matches_oracle(cuda, case):
    name, n, h, w, cin, cout, k, stride, mode, bn = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    x = rng.standard_normal((n, h, w, cin)).astype(F)
    wt = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(F)
    got, info, (ho, wo, pt, pl) = run_conv(x, wt, stride, mode, block_n=bn)
    want64 = ref64(x, wt, stride, pt, pl, ho, wo)
    want32 = L.conv2d(x, wt, stride, "SAME") if mode == "SAME" else L.conv2d_same(x, wt, stride)
    assert got.shape == want32.shape
    assert np.isfinite(got).all(), "%s: non-finite output (unwritten rows?) plan=%s" % (name, info)
    scale = np.abs(want64).max()
    e_gpu = np.abs(got - want64).max() / scale
    e_cpu = np.abs(want32 - want64).max() / scale
    print("\n[%s] plan=%s err_gpu_vs_f64=%.2e err_oracle_vs_f64=%.2e gpu_vs_oracle=%.2e" %
          (name, info, e_gpu, e_cpu, np.abs(got - want32).max() / scale))
    assert e_gpu < 4e-6, (name, e_gpu, info)


def test_conv_epilogue_bn_residual_relu(cuda):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 38, 50, 64)).astype(F)
    wt = (rng.standard_normal((1, 1, 64, 256)) * 0.2).astype(F)
    gamma, beta = rng.uniform(0.5, 1.5, 256).astype(F), rng.standard_normal(256).astype(F)
    mean, var = rng.standard_normal(256).astype(F), rng.uniform(0.5, 1.5, 256).astype(F)
    res = rng.standard_normal((1, 38, 50, 256)).astype(F)
    conv = L.conv2d(x, wt, 1, "SAME")
    bn, inv, shift = L.batch_norm(conv, gamma, beta, mean, var, 1e-5)
    want = L.relu(res + bn)
    got, info, _ = run_conv(x, wt, 1, "SAME", scale=inv, shift=shift, residual=res, act=1)
    err = np.abs(got - want).max()
    print("\n[epilogue] plan=%s max abs err=%.2e" % (info, err))
    assert err < 2e-5
    got6, _, _ = run_conv(x, wt, 1, "SAME", scale=None, shift=beta, act=2)
    assert np.abs(got6 - L.relu6(conv + beta)).max() < 2e-5


@pytest.mark.parametrize("shape", [("k3136_fc", 1, 1, 300, 3136, 128, 1), ("vgg_conv5", 1, 38, 50, 512, 512, 3),
                                   ("vgg_conv3", 1, 150, 200, 256, 256, 3), ("res_head_pw", 300, 7, 7, 2048, 512, 1),
                                   ("res_head_c3", 300, 7, 7, 512, 512, 3), ("res_b3_pw", 1, 38, 50, 1024, 256, 1)])
def test_accumulation_chunk_sweep(cuda, shape):
    """Accuracy and speed as a function of kb_per_chunk (k-blocks summed in TMEM before promotion)."""
    name, n, h, w, cin, cout, k = shape
    rng = np.random.default_rng(1)
    x = np.maximum(rng.standard_normal((n, h, w, cin)), 0).astype(F)      # post-ReLU-like
    wt = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(F)
    want64 = None
    for kpc in (2, 4, 8, 12, 16, 100000):
        for bn in ((0,) if kpc != 8 else (0, 64)):
            got, info, (ho, wo, pt, pl) = run_conv(x, wt, 1, "SAME", block_n=bn, kb_per_chunk=kpc, time_it=True)
            if want64 is None:
                want64 = ref64(x, wt, 1, pt, pl, ho, wo)
                want32 = L.conv2d(x, wt, 1, "SAME")
                print("\n[%s] oracle fp32 vs f64: %.2e" % (name, np.abs(want32 - want64).max() / np.abs(want64).max()))
            e = np.abs(got - want64).max() / np.abs(want64).max()
            print("[%s] kb_per_chunk=%d bn=%d grid=%dx%d tile=%dx%dx%d  err=%.2e  %.1f us  %.1f TFLOP/s" %
                  (name, kpc, info["block_n"], info["grid_m"], info["grid_n"], info["tile_n"], info["tile_h"], info["tile_w"],
                   e, info["us"], info["tflops"]))
            if kpc <= 8:
                assert e < 4e-6


@pytest.mark.parametrize("shape", [("b3_conv1", 1, 38, 50, 1024, 256, 1), ("b3_conv2", 1, 38, 50, 256, 256, 3), ("vgg_conv5", 1, 38, 50, 512, 512, 3),
                                   ("fc6_like", 1, 1, 300, 25088, 256, 1)])
def test_split_k(cuda, shape):
    """split-K (two-pass, deterministic) matches the unsplit kernel's accuracy, incl. BN + residual + ReLU epilogue."""
    name, n, h, w, cin, cout, k = shape
    rng = np.random.default_rng(3)
    x = np.maximum(rng.standard_normal((n, h, w, cin)), 0).astype(F)
    wt = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(F)
    scale = rng.uniform(0.5, 1.5, cout).astype(F); shift = rng.standard_normal(cout).astype(F)
    res = rng.standard_normal((n, h, w, cout)).astype(F)
    want = L.relu(L.conv2d(x, wt, 1, "SAME") * scale + shift + res)
    for sk in (1, 0, 2, 3, 8):
        got, info, _ = run_conv(x, wt, 1, "SAME", scale=scale, shift=shift, residual=res, act=1, split_k=sk, time_it=True)
        got2, _, _ = run_conv(x, wt, 1, "SAME", scale=scale, shift=shift, residual=res, act=1, split_k=sk)
        e = np.abs(got - want).max() / np.abs(want).max()
        print("\n[%s] split_k=%d -> splits=%d grid=%dx%d bn=%d  err vs oracle %.2e  %.1f us" %
              (name, sk, info["splits"], info["grid_m"], info["grid_n"], info["block_n"], e, info["us"]))
        assert e < 5e-6
        assert np.array_equal(got, got2), "split-K must be run-to-run deterministic"


def test_throughput_mode_f16x1(cuda):
    """FRCNN_CONV_F16X1 (opt-in, NOT fp32-grade): plain fp16 operands, one MMA per product, fp32 accumulation -- error of a few
    1e-4 of the output range, three orders above the default FP16x3 path on the same packed weights."""
    from tf_faster_rcnn_b200 import ops, _native as N
    rng = np.random.default_rng(12)
    for (n, h, w, cin, cout, k) in [(1, 38, 50, 256, 256, 3), (20, 7, 7, 512, 512, 1), (1, 1, 300, 3136, 128, 1)]:
        x = np.maximum(rng.standard_normal((n, h, w, cin)), 0).astype(np.float32)
        wt = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
        ho, wo, pt, pl = ops.conv_out_hw(h, w, k, 1, "SAME")
        errs = {}
        for impl in (N.CONV_F16X3, N.CONV_F16X1):
            pc = ops.PackedConv(wt, impl=impl)
            out = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.float32, device="cuda")
            ops.ConvPlan(torch.from_numpy(x).cuda(), pc, out, 1, pt, pl, 0, None).run()
            torch.cuda.synchronize()
            errs[impl] = out.cpu().numpy()
        scale = np.abs(errs[N.CONV_F16X3]).max()
        dev = np.abs(errs[N.CONV_F16X1] - errs[N.CONV_F16X3]).max() / scale
        assert not np.isnan(errs[N.CONV_F16X1]).any()
        assert 1e-6 < dev < 3e-3, dev
