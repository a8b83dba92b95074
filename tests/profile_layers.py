"""Per-launch timing table of one try-on step (GPU box).  Usage: python tests/profile_layers.py [clips] [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from shineon_virtual_tryon_b200 import ops  # noqa: E402
from shineon_virtual_tryon_b200.pipeline import TryOnPipeline  # noqa: E402


def main():
    # A/B switches: SHINEON_ACC_CHUNK (ops.ACC_CHUNK_KB), SHINEON_FUSE_STATS=0 (separate InstanceNorm statistics pass)
    if "SHINEON_ACC_CHUNK" in os.environ:
        ops.ACC_CHUNK_KB = int(os.environ["SHINEON_ACC_CHUNK"])
    if os.environ.get("SHINEON_FUSE_STATS") == "0":
        from shineon_virtual_tryon_b200.networks.cpvton.unet import UnetSkipConnectionBlock

        UnetSkipConnectionBlock.FUSE_STATS = False
    print(f"ACC_CHUNK_KB={ops.ACC_CHUNK_KB} FUSE_STATS={os.environ.get('SHINEON_FUSE_STATS', '1')}")
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
    dev = torch.device("cuda")
    warp, tom = bench.build_models()
    pipe = TryOnPipeline(warp.to(dev), tom.to(dev))
    pipe.set_precision(prec)
    frames = clips * 5
    a, c, p = (t.to(dev) for t in bench.synth_inputs(frames, 1))
    if os.environ.get("SHINEON_RAW") == "1":  # the uint8 path bench.py times (frame prep fused into the stems' operands)
        raw = [bench.synth_raw_frames(frames, 1, pinned=False)[k].to(dev) for k in TryOnPipeline.RAW_KEYS]
        prep = ops.FramePrep(256, 192, device=dev)
        call = lambda: pipe.run_raw(*raw, prep)
    else:
        call = lambda: pipe(a, c, p)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    prof = []
    ops.PROFILE = prof
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cuprof = os.environ.get("SHINEON_CUPROF") == "1"  # ncu --profile-from-start off: capture exactly this step
    if cuprof:
        torch.cuda.profiler.start()
    e0.record()
    call()
    e1.record()
    if cuprof:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    ops.PROFILE = None
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    print(f"step {total:.3f} ms for {frames} frames ({frames / total * 1e3:.0f} fps) precision {prec}")
    tconv = 0.0
    print(f"{'N':>4} {'HxW':>9} {'Cin':>5} {'Cout':>5} k s | {'ms':>8} {'TF/s':>8} {'GF':>8}")
    for rec in prof:  # (algorithmic FLOPs, start, end, shape[, issued FLOPs]): rows with a 5th field = low-res GEMM + gather
        fl, s, e, (N, H, W, Cin, cpad, Cout, k, st) = rec[:4]
        ms = s.elapsed_time(e)
        tconv += ms
        print(f"{N:4d} {H:4d}x{W:<4d} {Cin:5d} {Cout:5d} {k} {st} | {ms:8.3f} {fl / ms / 1e9:8.1f} {fl / 1e9:8.2f}"
              + (f"  (issued {rec[4] / 1e9:.2f} GF: low-res GEMM + gather)" if len(rec) > 4 else ""))
    print(f"conv total {tconv:.3f} ms = {tconv / total * 100:.1f}% of step; non-conv {total - tconv:.3f} ms")


if __name__ == "__main__":
    main()
