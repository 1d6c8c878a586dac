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
    tr = torch.zeros(64 * 8 + 192, dtype=torch.int64, device="cuda")
    N.check(N.lib().frcnn_conv_plan_set_trace(plan._h, C.c_void_p(tr.data_ptr())))
    plan.run(); torch.cuda.synchronize()
    N.check(N.lib().frcnn_conv_plan_set_trace(plan._h, C.c_void_p(0)))
    full = tr.cpu().numpy().astype(np.int64); t = full[:512].reshape(64, 8)
    nkb = min(64, k * k * cin // 32)
    t0 = t[0, 0]
    print("conv n=%d %dx%d cin=%d cout=%d k=%d plan=%s" % (n, h, w, cin, cout, k, plan.info()))
    print("kb | mma_seen  +3mma  +12mma  +commit | gap since previous commit")
    for i in range(8, min(nkb, 28)):
        print("%2d | %8d %6d | gap %6d" % (i, t[i, 4] - t0, t[i, 5] - t[i, 4], t[i, 4] - t[i - 1, 5]))
    nch = min(24, (nkb + kpc - 1) // kpc)
    print("chunk: mma_wait_begin mma_got_tmem_empty | epi_seen epi_drained (drain time)")
    for c in range(4, 10):
        print("  %2d %8d | %8d %8d (%5d)" % (c, full[576 + c] - t0, t[c, 7] - t0, full[512 + c] - t0, full[512 + c] - t[c, 7]))

import sys as _s
quiet = True
trace(300, 7, 7, 512, 512, 3, 128, 2)
