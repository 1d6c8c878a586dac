"""Debug aid (GPU box; needs the development build: FRCNN_LIB_VARIANT=wd python tools/trace_f16.py): clock64 timeline of CTA 0's pipeline hand-offs in conv_gemm_f16x3_kernel for its first 64 k-blocks.

columns per k-block: S0 splitter saw a_full | S1 converted (starts waiting ta_empty) | S2 got ta_empty | S3 arrived ta_full |
M4 issuer starts waiting | M5 operands ready | M6 MMAs issued + committed ; per chunk: E7 epilogue saw acc_full"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tf_faster_rcnn_b200 import ops, _native as N  # noqa: E402


def trace(n, h, w, cin, cout, k, kpc=4):
    rng = np.random.default_rng(0)
    x = torch.from_numpy(np.maximum(rng.standard_normal((n, h, w, cin)), 0).astype(np.float32)).cuda()
    wt = (rng.standard_normal((k, k, cin, cout)) * 0.05).astype(np.float32)
    pc = ops.PackedConv(wt)
    ho, wo, pt, pl = ops.conv_out_hw(h, w, k, 1, "SAME")
    out = torch.empty((n, ho, wo, cout), dtype=torch.float32, device="cuda")
    plan = ops.ConvPlan(x, pc, out, 1, pt, pl, 0, None, 0, kpc)
    for _ in range(3):
        plan.run()
    tr = torch.zeros(64 * 8 + 256, dtype=torch.int64, device="cuda")
    N.check(N.lib().frcnn_conv_plan_set_trace(plan._h, C.c_void_p(tr.data_ptr())))
    plan.run(); torch.cuda.synchronize()
    N.check(N.lib().frcnn_conv_plan_set_trace(plan._h, C.c_void_p(0)))
    t = tr.cpu().numpy().astype(np.int64)[:512].reshape(64, 8)
    t0 = t[0, 0]
    print("conv n=%d %dx%d cin=%d cout=%d k=%d plan=%s" % (n, h, w, cin, cout, k, plan.info()))
    print("kb |   S0     S1     S2     S3 |   M4     M5     M6 | S work  S wait_ta  st+arrive | M wait  M issue | M5-S3  S2-M6(kb-2)  period(M5)")
    for i in range(8, 40):
        r = t[i] - t0
        print("%2d | %6d %6d %6d %6d | %6d %6d %6d | %6d %8d %9d | %6d %7d | %5d %11d %10d" % (
            i, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[1] - r[0], r[2] - r[1], r[3] - r[2], r[5] - r[4], r[6] - r[5], r[5] - r[3],
            t[i, 2] - t[i - 2, 6], t[i, 5] - t[i - 1, 5]))
    print("chunk E7 (acc_full seen, cycles): ", [int(v - t0) for v in t[:10, 7]])


if __name__ == "__main__":
    trace(300, 7, 7, 512, 512, 3)
    trace(300, 7, 7, 1024, 2048, 1)
