"""Profiling aid: run ONE conv shape a few times (for `ncu -k regex:conv_gemm -s 3 -c 1 python tools/one_conv.py ...`)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tf_faster_rcnn_b200 import ops

n, h, w, cin, cout, k, res = [int(v) for v in sys.argv[1:8]]
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.standard_normal((n, h, w, cin)).astype(np.float32)).cuda()
pc = ops.PackedConv((rng.standard_normal((k, k, cin, cout)) * 0.05).astype(np.float32), np.ones(cout, np.float32), np.zeros(cout, np.float32))
out = torch.empty((n, h, w, cout), dtype=torch.float32, device="cuda")
r = torch.randn((n, h, w, cout), device="cuda") if res else None
plan = ops.ConvPlan(x, pc, out, 1, k // 2, k // 2, 1, r)
for _ in range(6):
    plan.run()
torch.cuda.synchronize()
print("done", plan.info())
