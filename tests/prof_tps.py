"""One fused TPS + grid_sample launch at batch B (ncu target).  usage: prof_tps.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gmm import TpsTables  # noqa: E402  (constants only)
from shineon_virtual_tryon_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H, W = 256, 192
t = TpsTables(H, W, 5)
dev = ops.TpsTablesDev(t.Li, t.P_X, t.P_Y, t.grid_X[0, :], t.grid_Y[:, 0], 5, "cuda")
theta = (torch.rand(B, 50, device="cuda") * 2 - 1) * 0.1
cloth = torch.rand(B, 3, H, W, device="cuda")
for _ in range(3):
    ops.tps_grid_sample(theta, dev, H, W, [(cloth, "border")])
torch.cuda.synchronize()
