"""Roofline measurements of the memory-bound ops at the BASELINE config sizes (GPU box).
Prints one JSON object per op: algorithmic bytes, CUDA-event time, GB/s, fraction of measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gmm import TpsTables  # noqa: E402  (constants only)
from shineon_virtual_tryon_b200 import ops  # noqa: E402

PEAK = 6584.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, iters=20, flush=None, reps=10):
    """Median CUDA-event time of one launch.  Small configs (flush given): the 126 MB L2 is evicted before every
    launch, one launch per event pair (launch latency included: these sizes are latency-bound anyway).  Large configs
    (data > L2): `reps` back-to-back launches per event pair, so the ~5-10 us launch gap of a lone 50 us kernel does not
    pollute a bandwidth figure; every launch still streams from HBM."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # 512 MB write: evicts the 126 MB L2
        n = 1 if flush is not None else reps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, nbytes, ms, extra=None):
    gbs = nbytes / ms / 1e6
    d = {"op": name, "algorithmic_MB": round(nbytes / 1e6, 2), "ms": round(ms, 4), "GB/s": round(gbs, 1),
         "frac_of_measured_hbm": round(gbs / PEAK, 3)}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    H, W = 256, 192
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    t = TpsTables(H, W, 5)
    dev = ops.TpsTablesDev(t.Li, t.P_X, t.P_Y, t.grid_X[0, :], t.grid_Y[:, 0], 5, "cuda")
    for B in (8, 512):
        theta = (torch.rand(B, 50, device="cuda") * 2 - 1) * 0.1
        cloth = torch.rand(B, 3, H, W, device="cuda")
        fl = flush if B == 8 else None  # B=512 streams 600 MB: larger than L2 by itself
        ms = timeit(lambda: ops.tps_grid_sample(theta, dev, H, W, [(cloth, "border")]), flush=fl)
        report(f"tps_grid_sample fused B={B}", B * 6 * H * W * 4, ms, {"l2": "flushed" if fl is not None else "data > L2"})
        grid = ops.tps_grid(theta, dev, H, W)
        ms = timeit(lambda: ops.grid_sample(cloth, grid, "border"), flush=fl)
        report(f"grid_sample (grid materialised) B={B}", B * 8 * H * W * 4, ms)
    for B in (16, 256):
        img = torch.rand(B, 3, H, W, device="cuda")
        flow = torch.randn(B, 2, H, W, device="cuda") * 4
        fl = flush if B == 16 else None
        ms = timeit(lambda: ops.resample2d_fwd(img, flow), flush=fl)
        report(f"resample2d B={B} (i.i.d. N(0,4px) flow, SURVEY config 4)", B * 8 * H * W * 4, ms)
        if B == 256:  # a smooth flow field (what FlowNet2 produces): neighbouring pixels sample neighbouring texels
            sm = torch.nn.functional.interpolate(torch.randn(B, 2, H // 16, W // 16, device="cuda") * 4, size=(H, W),
                                                 mode="bilinear", align_corners=False).contiguous()
            ms = timeit(lambda: ops.resample2d_fwd(img, sm))
            report(f"resample2d B={B} (smooth flow)", B * 8 * H * W * 4, ms)
        ms = timeit(lambda: ops.channelnorm_fwd(img), flush=fl)
        report(f"channelnorm C=3 B={B}", B * 4 * H * W * 4, ms)
    for B in (16, 64):
        f1 = torch.randn(B, 256, 32, 24, device="cuda")
        f2 = torch.randn(B, 256, 32, 24, device="cuda")
        ms = timeit(lambda: ops.correlation_fwd(f1, f2, 20, 1, 20, 1, 2), flush=flush if B == 16 else None)
        report(f"correlation (FlowNetC cfg) B={B}", B * (2 * 256 + 441) * 32 * 24 * 4, ms,
               {"GFLOP/s": round(B * 0.173e9 / ms / 1e6, 1)})


if __name__ == "__main__":
    main()
