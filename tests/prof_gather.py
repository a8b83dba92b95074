"""One launch of each memory-bound op at a large batch (for ncu --set full captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gmm import TpsTables  # noqa: E402
from shineon_virtual_tryon_b200 import ops  # noqa: E402

H, W, B = 256, 192, 256
t = TpsTables(H, W, 5)
dev = ops.TpsTablesDev(t.Li, t.P_X, t.P_Y, t.grid_X[0, :], t.grid_Y[:, 0], 5, "cuda")
theta = (torch.rand(B, 50, device="cuda") * 2 - 1) * 0.1
cloth = torch.rand(B, 3, H, W, device="cuda")
flow = torch.randn(B, 2, H, W, device="cuda") * 4
for _ in range(2):
    ops.tps_grid_sample(theta, dev, H, W, [(cloth, "border")])
    grid = ops.tps_grid(theta, dev, H, W)
    ops.grid_sample(cloth, grid, "border")
    ops.resample2d_fwd(cloth, flow)
    ops.channelnorm_fwd(cloth)
torch.cuda.synchronize()
f1 = torch.randn(64, 256, 32, 24, device="cuda")
f2 = torch.randn(64, 256, 32, 24, device="cuda")
for _ in range(2):
    ops.correlation_fwd(f1, f2, 20, 1, 20, 1, 2)
torch.cuda.synchronize()
