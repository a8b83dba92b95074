"""Bring-up diagnostic for the MN-major wgrad kernel: both descriptor variants, a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shineon_virtual_tryon_b200 import ops

for (N, H, W, Cin, Cout, k, s, p) in [(2, 16, 12, 64, 64, 3, 1, 1), (1, 8, 8, 64, 64, 1, 1, 0), (2, 16, 12, 128, 192, 4, 2, 1)]:
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(N, Cin, H, W, generator=gen)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    g = torch.randn(N, Cout, Ho, Wo, generator=gen)
    want = torch.nn.grad.conv2d_weight(x, (Cout, Cin, k, k), g, stride=s, padding=p)
    for prec in ("bf16", "bf16x3"):
        xp = ops.nchw_to_planes(x.cuda(), prec=prec)
        gp = ops.nchw_to_planes(g.cuda(), prec=prec)
        for variant in (0, 1):
            gw = torch.zeros(Cout, Cin, k, k, device="cuda")
            try:
                ops.conv2d_wgrad(gp, xp, gw, Cout=Cout, Cin=Cin, kh=k, kw=k, stride=s, pad=p, desc_variant=variant)
                torch.cuda.synchronize()
                err = ((gw.cpu() - want).abs().max() / want.abs().max()).item()
            except Exception as e:  # noqa: BLE001
                err = f"EXC {e}"
            print(f"case {(N,H,W,Cin,Cout,k,s,p)} {prec} variant {variant}: rel err {err}", flush=True)
