"""Per-launch timing table of one FlowNet2 forward (GPU box).  Usage: python tests/profile_flownet.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import _lib, ops  # noqa: E402
from shineon_virtual_tryon_b200.models.flownet import FlowNet  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.manual_seed(420)
    net = FlowNet()
    for m in net.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            torch.nn.init.xavier_uniform_(m.weight)
    net = net.cuda().eval()
    net.flowNet.parallel_sd = False  # one stream: per-launch events must not include a neighbour's work
    g = torch.Generator().manual_seed(1)
    im1, im2 = torch.rand(B, 3, 256, 192, generator=g).cuda(), torch.rand(B, 3, 256, 192, generator=g).cuda()
    with torch.no_grad():
        for _ in range(3):
            net(im1, im2)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        prof = []
        ops.PROFILE = prof
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cuprof = os.environ.get("SHINEON_CUPROF") == "1"  # ncu --profile-from-start off: capture exactly this forward
        if cuprof:
            torch.cuda.profiler.start()
        e0.record()
        net(im1, im2)
        e1.record()
        if cuprof:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        ops.PROFILE = None
        torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    print(f"FlowNet2 forward B={B}: {total:.3f} ms, {_lib.launch_count() - l0} shineon launches, {len(prof)} tensor-core launches")
    rows = []
    for rec in prof:
        fl, s, e, (N, H, W, Cin, cpad, Cout, k, st) = rec[:4]
        rows.append((s.elapsed_time(e), fl, N, H, W, Cin, cpad, Cout, k, st))
    tconv = sum(r[0] for r in rows)
    print(f"conv total {tconv:.3f} ms = {tconv / total * 100:.1f}% of the forward; the rest {total - tconv:.3f} ms")
    print(f"{'ms':>8} {'TF/s':>7} {'GF':>7} |  N  HxW  Cin(cpad) Cout k s")
    for ms, fl, N, H, W, Cin, cpad, Cout, k, st in sorted(rows, reverse=True)[:45]:
        print(f"{ms:8.3f} {fl / ms / 1e9:7.1f} {fl / 1e9:7.2f} | {N:3d} {H:4d}x{W:<4d} {Cin:4d}({cpad:4d}) {Cout:4d} {k} {st}")


if __name__ == "__main__":
    main()
