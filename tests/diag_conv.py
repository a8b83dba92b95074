"""Diagnostic ladder for the tcgen05 conv kernel (run on the GPU box; not a pytest).

Walks from the simplest possible GEMM (one tile, one K-block) to the real layer shapes and prints, for
each, the error of the tcgen05 kernel against the CUDA-core direct kernel and the CPU oracle, plus the
structure of the error when it fails (which rows / columns / K-blocks are off).
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shineon_virtual_tryon_b200 import ops  # noqa: E402


def run(name, N, H, W, Cin, Cout, k, s, p, split, tile_n=0, stages=0, check_cpu=True):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    prec = 'fp16x3' if split else 'bf16'
    xp = ops.nchw_to_planes(x.cuda(), prec=prec)
    pc = ops.PackedConv(w.cuda(), b.cuda(), stride=s, pad=p, prec=prec)
    try:
        yd, _ = ops.conv2d(xp, pc, want_f32=True, direct=True)
        torch.cuda.synchronize()
        t0 = time.time()
        y, _ = ops.conv2d(xp, pc, want_f32=True, tile_n=tile_n, stages=stages)
        torch.cuda.synchronize()
        dt = time.time() - t0
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] EXCEPTION {type(e).__name__}: {e}")
        return False
    err = (y - yd).abs()
    mx = err.max().item()
    ok = mx < 1e-3 and bool(torch.isfinite(y).all())
    msg = f"[{name}] N{N} {H}x{W} Cin{Cin} Cout{Cout} k{k} s{s} p{p} split={split} bn={tile_n} st={stages}: igemm-vs-direct max {mx:.3e}"
    if check_cpu:
        xs, ws = (x, w) if split else (x.bfloat16().float(), w.bfloat16().float())
        want = F.conv2d(xs, ws, b, stride=s, padding=p).permute(0, 2, 3, 1)
        msg += f" | direct-vs-cpu {(yd.cpu() - want).abs().max().item():.3e} | igemm-vs-cpu {(y.cpu() - want).abs().max().item():.3e}"
    print(msg + f" | {'OK' if ok else 'FAIL'} ({dt * 1e3:.1f} ms)")
    if not ok:
        e2 = err.reshape(-1, Cout)
        bad_rows = (e2.max(1).values > 1e-3).nonzero().flatten()
        bad_cols = (e2.max(0).values > 1e-3).nonzero().flatten()
        print(f"    bad pixels {bad_rows.numel()}/{e2.shape[0]} first {bad_rows[:16].tolist()}")
        print(f"    bad channels {bad_cols.numel()}/{Cout} first {bad_cols[:16].tolist()}")
        print("    got ", y.reshape(-1, Cout)[:2, :8].tolist())
        print("    want", yd.reshape(-1, Cout)[:2, :8].tolist())
        ratio = (y.reshape(-1, Cout)[:4, :4] / yd.reshape(-1, Cout)[:4, :4])
        print("    ratio", ratio.tolist())
    return ok


def main():
    print(torch.cuda.get_device_name(0))
    ladder = [
        ("gemm-1kb", 1, 8, 16, 64, 64, 1, 1, 0, False),
        ("gemm-1kb-split", 1, 8, 16, 64, 64, 1, 1, 0, True),
        ("gemm-4kb", 1, 8, 16, 256, 64, 1, 1, 0, False),
        ("gemm-bn128", 1, 8, 16, 128, 128, 1, 1, 0, False),
        ("gemm-bn16", 1, 8, 16, 128, 16, 1, 1, 0, False),
        ("gemm-bn32", 1, 8, 16, 128, 32, 1, 1, 0, False),
        ("gemm-mtiles", 2, 16, 16, 64, 64, 1, 1, 0, False),
        ("conv3", 1, 8, 16, 64, 64, 3, 1, 1, False),
        ("conv3-split", 2, 16, 12, 128, 128, 3, 1, 1, True),
        ("conv4s2", 2, 16, 16, 64, 64, 4, 2, 1, False),
        ("conv4s2-split", 3, 32, 24, 22, 64, 4, 2, 1, True),
        ("tiny-spatial", 5, 4, 3, 512, 512, 3, 1, 1, True),
        ("final-layer", 1, 64, 48, 128, 4, 3, 1, 1, True),
        ("k7s2", 2, 32, 32, 3, 64, 7, 2, 3, True),
    ]
    res = [run(*c) for c in ladder]
    # BN=256 and stage overrides
    res.append(run("bn256", 2, 16, 16, 128, 256, 3, 1, 1, True, 256, 0))
    res.append(run("stages1", 2, 16, 16, 128, 128, 3, 1, 1, True, 0, 1))
    res.append(run("stages2", 2, 16, 16, 128, 128, 3, 1, 1, True, 0, 2))
    print(f"diag: {sum(res)}/{len(res)} passed")


if __name__ == "__main__":
    main()
