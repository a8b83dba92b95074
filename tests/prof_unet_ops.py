"""A few launches of the U-Net's non-GEMM hot kernels at the bench shapes (for ncu --set full captures and quick timing):
upconv3x3_gather (decoder level 64x48 -> 128x96, 256 -> 64 channels), sagan_attention (N = 192, C = 512), 80 images."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 80
g = torch.Generator(device="cuda").manual_seed(0)


def timed(name, fn, nbytes=None, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.4f} ms" + (f"  {nbytes / ms / 1e6:.0f} GB/s algorithmic" if nbytes else ""), flush=True)


# decoder level: low-res [N,64,48,256] -> [N,128,96,64]
w = torch.randn(64, 256, 3, 3, device="cuda", generator=g) * 0.02
b = torch.randn(64, device="cuda", generator=g) * 0.1
up = ops.UpsampledConv3x3(w, b)
x = ops.nchw_to_planes(torch.randn(N, 256, 64, 48, device="cuda", generator=g))
t, _ = ops.conv2d(x, up.pc, want_f32=True)
y = torch.empty(N, 128, 96, 64, device="cuda")
from shineon_virtual_tryon_b200 import _lib  # noqa: E402


def gather():
    ops.check(_lib.load().shineon_upconv3x3_gather(ops._p(t), ops._p(b), ops._p(y), N, 64, 48, 64, t.shape[-1], ops._stream()),
              "gather")


timed("tap-stacked GEMM 256->576 @64x48", lambda: ops.conv2d(x, up.pc, out_f32=t))
timed("upconv3x3_gather Cout=64 @64x48->128x96", gather, nbytes=t.numel() * 4 + y.numel() * 4)
# final level: low-res [N,128,96,128] -> [N,256,192,4]
w4 = torch.randn(4, 128, 3, 3, device="cuda", generator=g) * 0.02
up4 = ops.UpsampledConv3x3(w4, None)
x4 = ops.nchw_to_planes(torch.randn(N, 128, 128, 96, device="cuda", generator=g))
t4, _ = ops.conv2d(x4, up4.pc, want_f32=True)
y4 = torch.empty(N, 256, 192, 4, device="cuda")
timed("tap-stacked GEMM 128->36 @128x96", lambda: ops.conv2d(x4, up4.pc, out_f32=t4))
timed("upconv3x3_gather Cout=4 @128x96->256x192",
      lambda: ops.check(_lib.load().shineon_upconv3x3_gather(ops._p(t4), ops._p(None), ops._p(y4), N, 128, 96, 4, t4.shape[-1],
                                                             ops._stream()), "gather"),
      nbytes=t4.numel() * 4 + y4.numel() * 4)
# attention at the three sizes of the ShineOn U-Net
for hw in ((16, 12), (8, 6), (4, 3)):
    HW = hw[0] * hw[1]
    qkv = torch.randn(N, hw[0], hw[1], 640, device="cuda", generator=g)
    xx = torch.randn(N, hw[0], hw[1], 512, device="cuda", generator=g)
    gm = torch.full((1,), 0.7, device="cuda")
    timed(f"sagan_attention N={HW} C=512", lambda: ops.sagan_attention(qkv, xx, gm, 64, act="gelu", want_f32=False, want_planes=True))
# instance norm + GELU -> planes on the largest level
c = torch.randn(N, 128, 96, 64, device="cuda", generator=g)
timed("instnorm_act [N,128,96,64] -> planes", lambda: ops.instnorm_act(c, act="gelu", want_planes=True),
      nbytes=c.numel() * 12)
# first-layer layout conversions (NCHW f32 -> space-to-depth planes): GMM person 22 ch, U-Net person+cloth 7+3 ch
for c0, c1 in ((22, 0), (7, 3)):
    x0 = torch.randn(N, c0, 256, 192, device="cuda", generator=g)
    x1 = torch.randn(N, c1, 256, 192, device="cuda", generator=g) if c1 else None
    wt = torch.randn(64, c0 + c1, 4, 4, device="cuda", generator=g) * 0.02
    s2d = ops.S2dConv(wt, None)
    nb = N * (c0 + c1) * 256 * 192 * 4 + N * 129 * 97 * ops.cpad64(4 * (c0 + c1)) * 4
    timed(f"nchw_s2d_planes C={c0 + c1}", lambda: s2d.prepare(x0, x1), nbytes=nb)
# GMM correlation
fa = torch.randn(N, 16, 12, 512, device="cuda", generator=g)
fb = torch.randn(N, 16, 12, 512, device="cuda", generator=g)
timed("l2norm_correlation 16x12x512", lambda: ops.l2norm_correlation(fa, fb, want_f32=False, want_planes=True))
