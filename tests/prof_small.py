"""One launch each of sagan_attention (N=192, C=512) and instnorm_act ([N,128,96,64] -> planes) for ncu --set full."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402

N = 80
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(N, 16, 12, 640, device="cuda", generator=g)
xx = torch.randn(N, 16, 12, 512, device="cuda", generator=g)
gm = torch.full((1,), 0.7, device="cuda")
c = torch.randn(N, 128, 96, 64, device="cuda", generator=g)
for _ in range(2):
    ops.sagan_attention(qkv, xx, gm, 64, act="gelu", want_f32=False, want_planes=True)
    ops.instnorm_act(c, act="gelu", want_planes=True)
torch.cuda.synchronize()
