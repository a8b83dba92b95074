"""One launch each of the kernels added late in round 2, at the BASELINE configs[3] shapes, between cudaProfilerStart/Stop
(for `ncu --set full --profile-from-start off`): the FlowNetC cost volume (per-image tcgen05 GEMM + the gather that writes
LeakyReLU(cost volume) as NHWC planes into a concat window) and two one-launch transposed convs (deconv5 of FlowNetS:
16 x 4x3 x 1024 -> 512, and FlowNetFusion's deconv0: 16 x 128x96 x 162 -> 16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402
from shineon_virtual_tryon_b200.networks.deconv import PackedDeconv4x4s2  # noqa: E402

B = 16
g = torch.Generator().manual_seed(3)
f1 = ops.nchw_to_planes(torch.randn(B, 256, 32, 24, generator=g).cuda())
f2 = ops.nchw_to_planes(torch.randn(B, 256, 32, 24, generator=g).cuda())
cat = ops.Planes(B, 32, 24, 64 + 448, device="cuda", cpad=512)
cat.hi.zero_()
cat.lo.zero_()
x5 = ops.nchw_to_planes(torch.randn(B, 1024, 4, 3, generator=g).cuda())
d5 = PackedDeconv4x4s2((torch.randn(1024, 512, 4, 4, generator=g) * 0.02).cuda(), torch.zeros(512).cuda())
x0 = ops.nchw_to_planes(torch.randn(B, 162, 128, 96, generator=g).cuda())
d0 = PackedDeconv4x4s2((torch.randn(162, 16, 4, 4, generator=g) * 0.02).cuda(), torch.zeros(16).cuda())


def run():
    ops.correlation_planes(f1, f2, 256, 20, 20, 2, out_planes=cat.window(64, 441), act="leaky", act_param=0.1)
    d5(x5, post_act="leaky", act_param=0.1, want_planes=True)
    d0(x0, post_act="leaky", act_param=0.1, want_planes=True)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
