"""Accuracy of the tensor-core accumulation vs K (run on the GPU box: python tests/diag_accum.py).

tcgen05.mma adds into its fp32 TMEM accumulator with truncation, so one accumulation chain over the whole K loses
~(MMA count) x 2^-25 relative; conv_igemm cuts K into chunks of 16 K-blocks (1024 K) whose partial sums are added in
registers (round-to-nearest).  This prints the error of a 3x3 conv against an fp64 reference for both settings."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from shineon_virtual_tryon_b200 import ops

    torch.manual_seed(0)
    print("| Cin | K | chain | max rel-to-rms err | rms rel err | mean signed err / rms |")
    print("|---|---|---|---|---|---|")
    for cin in (256, 1024, 2688):
        N, H, W, cout = 2, 16, 12, 128
        x = torch.randn(N, cin, H, W)
        w = torch.randn(cout, cin, 3, 3) * 0.02
        ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
        rms = ref.pow(2).mean().sqrt().item()
        xp = ops.nchw_to_planes(x.cuda())
        for chunk, label in ((-1, "whole K (round 1)"), (0, "1024-K chunks")):
            ops.ACC_CHUNK_KB = chunk
            pc = ops.PackedConv(w.cuda(), None, stride=1, pad=1)
            y, _ = ops.conv2d(xp, pc, want_f32=True)
            err = (y.permute(0, 3, 1, 2).double().cpu() - ref)
            print(f"| {cin} | {9 * cin} | {label} | {err.abs().max().item() / rms:.2e} | {err.pow(2).mean().sqrt().item() / rms:.2e} | "
                  f"{(err * ref.sign()).mean().item() / rms:+.2e} |")
        ops.ACC_CHUNK_KB = 0
    # fp32 reference points: torch CUDA conv with TF32 off (IEEE fp32 FMA chains)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for cin in (2688,):
        x = torch.randn(2, cin, 16, 12)
        w = torch.randn(128, cin, 3, 3) * 0.02
        ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
        rms = ref.pow(2).mean().sqrt().item()
        y = torch.nn.functional.conv2d(x.cuda(), w.cuda(), padding=1).double().cpu()
        err = y - ref
        print(f"| {cin} | {9 * cin} | torch fp32 (cuDNN, TF32 off), for scale | {err.abs().max().item() / rms:.2e} | "
              f"{err.pow(2).mean().sqrt().item() / rms:.2e} | {(err * ref.sign()).mean().item() / rms:+.2e} |")


if __name__ == "__main__":
    main()
