"""One small invocation of the hot path on cuda:0, checked against the oracle (called by
__graft_entry__.smoke() only; the oracle import lives here because smoke is a checker)."""


def run(np, torch):
    from oracle import layers as L, nms as ONMS
    from . import ops, _native as N
    F = np.float32
    rng = np.random.default_rng(0)
    # dense: 3x3 conv + bias + ReLU on tcgen05 (3xTF32)
    x = rng.standard_normal((1, 38, 50, 64)).astype(F)
    w = (rng.standard_normal((3, 3, 64, 128)) * 0.04).astype(F)
    b = rng.standard_normal(128).astype(F)
    want = L.relu(L.conv2d(x, w, 1, "SAME") + b)
    pc = ops.PackedConv(w, None, b)
    out = torch.empty((1, 38, 50, 128), dtype=torch.float32, device="cuda")
    plan = ops.ConvPlan(torch.from_numpy(x).cuda(), pc, out, 1, 1, 1, N.ACT_RELU)
    plan.run()
    torch.cuda.synchronize()
    err = float(np.abs(out.cpu().numpy() - want).max())
    assert err < 2e-5, "conv parity %g" % err
    # NMS through the `_nms`-compatible entry
    n = 500
    xy = rng.uniform(0, 400, (n, 2)); wh = rng.uniform(10, 120, (n, 2))
    dets = np.hstack([xy, xy + wh, rng.random((n, 1))]).astype(F)
    order = ONMS.argsort_desc(dets[:, 4])
    got = order[ops.nms_host(dets[order], float(ONMS.thresh_f32(0.3, True)), N.NMS_MODE_CPU_NMS)]
    assert np.array_equal(got, ONMS.nms_plus1_c(dets, 0.3, True)), "nms parity"
    return err
