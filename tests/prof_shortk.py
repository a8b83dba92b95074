"""One short-K conv launch with plane outputs (ncu target): the SAMS mlp_shared GEMM / the try-on stems' regime."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402

N, H, W, K, Cout = 8, 256, 192, 36, int(sys.argv[1]) if len(sys.argv) > 1 else 384
act = sys.argv[2] if len(sys.argv) > 2 else "gelu"
gen = torch.Generator().manual_seed(0)
x = ops.nchw_to_planes(torch.randn(N, K, H, W, generator=gen).cuda())
w = torch.randn(Cout, K, 1, 1, generator=gen).cuda() * 0.1
pc = ops.PackedConv(w, torch.zeros(Cout).cuda(), stride=1, pad=0)
for _ in range(3):
    ops.conv2d(x, pc, post_act=None if act == "none" else act, want_planes=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv2d(x, pc, post_act=None if act == "none" else act, want_planes=True)
e1.record()
torch.cuda.synchronize()
print(f"Cout={Cout} act={act}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch")
