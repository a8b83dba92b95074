"""A/B of the split-K plan on single conv layers, timed as CUDA-graph replays of 20 back-to-back launches (GPU box).
usage: python tests/prof_splitk.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402

SHAPES = [  # N, H, W, Cin, Cout, k, s, p  (try-on step at 80 frames, FlowNet2 at 16 pairs)
    (80, 8, 6, 512, 256, 4, 2, 1), (80, 8, 6, 512, 512, 4, 2, 1), (80, 16, 12, 512, 512, 4, 2, 1), (80, 4, 3, 256, 128, 3, 1, 1),
    (80, 16, 12, 192, 512, 4, 2, 1), (16, 4, 3, 1024, 1024, 3, 1, 1), (16, 8, 6, 1026, 512, 3, 1, 1), (16, 16, 12, 512, 512, 3, 1, 1),
    (16, 16, 12, 770, 256, 3, 1, 1), (16, 8, 6, 512, 1024, 3, 2, 1), (16, 32, 24, 473, 256, 3, 1, 1)]


def bench(x, pc, reps=20):
    for _ in range(2):
        ops.conv2d(x, pc, want_f32=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            ops.conv2d(x, pc, want_f32=True)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1e3)
    return sorted(ts)[2]


def main():
    gen = torch.Generator().manual_seed(0)
    print(f"{'shape':>40} | single-pass us | split-K us")
    for N, H, W, Cin, Cout, k, s, p in SHAPES:
        x = ops.nchw_to_planes(torch.randn(N, Cin, H, W, generator=gen).cuda())
        w = torch.randn(Cout, Cin, k, k, generator=gen).cuda() * 0.02
        res = []
        for split in (False, True):
            ops.SPLIT_K = split
            pc = ops.PackedConv(w, None, stride=s, pad=p)
            res.append(bench(x, pc))
            used = bool(getattr(pc, "_sk_ws", None))
        ops.SPLIT_K = False
        print(f"{str((N, H, W, Cin, Cout, k, s)):>40} | {res[0]:10.1f} | {res[1]:10.1f} {'(split)' if used else '(no split planned)'}")


if __name__ == "__main__":
    main()
