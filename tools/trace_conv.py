"""Debug aid (GPU box): dump the pipeline hand-off timeline of CTA(0,0) of one conv launch."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), ".."))
from tf_faster_rcnn_b200 import ops, _native as N

def trace(n, h, w, cin, cout, k, bn, kpc=2):
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((n, h, w, cin)).astype(np.float32)).cuda()
    wt = (rng.standard_normal((k, k, cin, cout)) * 0.05).astype(np.float32)
    pc = ops.PackedConv(wt)
    ho, wo, pt, pl = ops.conv_out_hw(h, w, k, 1, "SAME")
    out = torch.empty((n, ho, wo, cout), dtype=torch.float32, device="cuda")
    plan = ops.ConvPlan(x, pc, out, 1, pt, pl, 0, None, bn, kpc)
    for _ in range(3): plan.run()
    tr = torch.zeros(64 * 8 + 256, dtype=torch.int64, device="cuda")
    N.check(N.lib().frcnn_conv_plan_set_trace(plan._h, C.c_void_p(tr.data_ptr())))
    plan.run(); torch.cuda.synchronize()
    N.check(N.lib().frcnn_conv_plan_set_trace(plan._h, C.c_void_p(0)))
    full = tr.cpu().numpy().astype(np.int64); t = full[:512].reshape(64, 8)
    nkb = min(64, k * k * cin // 32)
    t0 = t[0, 0]
    print("conv n=%d %dx%d cin=%d cout=%d k=%d plan=%s" % (n, h, w, cin, cout, k, plan.info()))
    print("kb | mma_seen  +3mma  +12mma  +commit | gap since previous commit")
    for i in range(0, 0):
        print("%2d | %8d %6d | gap %6d" % (i, t[i, 4] - t0, t[i, 5] - t[i, 4], t[i, 4] - t[i - 1, 5]))
    nk_t = min(nkb, 64) - 1
    print("CTA timeline (cycles from kernel entry): prologue done %d | first TMA issued %d | first MMA seen %d | last traced MMA commit %d (kb %d) | "
          "epilogue start %d | stores done %d | after dealloc %d" % (full[701] - full[700], t[0, 1] - full[700], t[0, 4] - full[700], t[nk_t, 5] - full[700], nk_t,
           full[702] - full[700], full[703] - full[700], full[704] - full[700]))
    nch = min(24, (nkb + kpc - 1) // kpc)
    print("chunk: mma_wait_begin mma_got_tmem_empty | epi_seen epi_drained (drain time)")
    for c in range(0, 0):
        print("  %2d %8d | %8d %8d (%5d)" % (c, full[576 + c] - t0, t[c, 7] - t0, full[512 + c] - t0, full[512 + c] - t[c, 7]))


def timed(n, h, w, cin, cout, k, res=False, reps=20):
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((n, h, w, cin)).astype(np.float32)).cuda()
    pc = ops.PackedConv((rng.standard_normal((k, k, cin, cout)) * 0.05).astype(np.float32), np.ones(cout, np.float32), np.zeros(cout, np.float32))
    out = torch.empty((n, h, w, cout), dtype=torch.float32, device="cuda")
    r = torch.randn((n, h, w, cout), device="cuda") if res else None
    plan = ops.ConvPlan(x, pc, out, 1, k // 2, k // 2, 1, r)
    for _ in range(3): plan.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.run()
    e1.record(); torch.cuda.synchronize()
    print("timed n=%d cin=%d cout=%d k=%d res=%s: %.1f us  %s" % (n, cin, cout, k, res, e0.elapsed_time(e1) * 1000 / reps, plan.info()))

if len(sys.argv) > 1 and sys.argv[1] == "bench":
    a = [int(v) for v in sys.argv[2:9]]
    timed(a[0], a[1], a[2], a[3], a[4], a[5], res=bool(a[6]))
    sys.exit(0)
trace(300, 7, 7, 512, 2048, 1, 128, 8)
trace(300, 7, 7, 512, 512, 3, 128, 8)
for nroi in (37, 74, 300):      # 37 rois*49 = 1813 rows -> 15 m-tiles x 16 = 240 CTAs ; 74 -> 29 x 16 = 464
    timed(nroi, 7, 7, 512, 2048, 1, res=True)
    timed(nroi, 7, 7, 512, 2048, 1, res=False)
timed(9, 7, 7, 512, 2048, 1, res=True)       # 4 m-tiles x 16 = 64 CTAs: a single partial wave
timed(9, 7, 7, 512, 512, 1, res=False)
