"""Runs one conv layer shape a few times (for ncu captures).  usage: prof_one_conv.py N H W Cin Cout k s p [prec]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402

N, H, W, Cin, Cout, k, s, p = [int(v) for v in sys.argv[1:9]]
prec = sys.argv[9] if len(sys.argv) > 9 else "fp16x3"
planes_out = len(sys.argv) > 10 and sys.argv[10] == "planes"
x = ops.Planes(N, H, W, Cin, prec=prec)
x.hi.normal_()
if x.lo is not None:
    x.lo.normal_(std=1e-3)
w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.02
pc = ops.PackedConv(w, torch.zeros(Cout, device="cuda"), stride=s, pad=p, prec=prec)
for _ in range(4):
    y = ops.conv2d(x, pc, want_f32=not planes_out, want_planes=planes_out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv2d(x, pc, want_f32=not planes_out, want_planes=planes_out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
print(f"{ms:.3f} ms  {2.0 * N * Ho * Wo * Cout * k * k * Cin / ms / 1e9:.1f} TF/s")
