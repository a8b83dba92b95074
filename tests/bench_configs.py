"""Timings of the BASELINE.json configs that bench.py does not cover, on one GPU, CUDA events, median of
`iters` after 3 warm-up calls.  One JSON line per config:
  configs[0]  UnetMaskModel forward, 1 frame (latency)                         + the CPU oracle port's time for the same frame
  configs[1]  WarpModel GMM + TPS forward, batch 8 (conv time vs 74.4 GFLOP)   + batch 80
  configs[3]  FlowNet2 two-frame forward, batch 16 (49.56 GFLOP / pair)
(configs[2] = bench.py, configs[4] = bench.py --workload train; the memory-bound ops of configs[1]/[3] = tests/bench_ops.py)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

PEAK_TF = 1457.6
try:
    PEAK_TF = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda", 0)
    warp, tom = bench.build_models()
    warp, tom = warp.to(dev), tom.to(dev)
    g = torch.Generator().manual_seed(420)
    # configs[0]
    person, cloth = torch.randn(1, 7, 256, 192, generator=g).to(dev), torch.randn(1, 3, 256, 192, generator=g).to(dev)
    with torch.no_grad():
        ms = timeit(lambda: tom(person, cloth))
    line = {"config": "configs[0]: UnetMaskModel forward, 1 frame 256x192 (self-attn, GELU), fp16x3", "ms": round(ms, 4),
            "frames_per_s": round(1e3 / ms, 1), "algorithmic_gflop": 16.731, "tflops": round(16.731 / ms, 2)}
    try:
        from oracle import unet as ounet

        sd = {k: v.detach().cpu() for k, v in tom.state_dict().items()}
        pc, cc = person.cpu(), cloth.cpu()
        kw = dict(n_frames=1, flow_warp=False, num_downs=6, num_attention=2, use_self_attn=True, act="gelu")
        with torch.no_grad():
            ounet.tom_forward(sd, pc, cc, **kw)
            t0 = time.perf_counter()
            for _ in range(5):
                ounet.tom_forward(sd, pc, cc, **kw)
        line["cpu_oracle_port_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 2)
        line["cpu_threads"] = torch.get_num_threads()
    except Exception as e:  # noqa: BLE001  (the CPU column is informative only)
        line["cpu_oracle_port_ms"] = None
        line["cpu_note"] = str(e)[:80]
    print(json.dumps(line), flush=True)
    # configs[1]
    for B in (8, 80):
        a = torch.randn(B, 22, 256, 192, generator=g).to(dev)
        c = (torch.rand(B, 3, 256, 192, generator=g) * 2 - 1).to(dev)
        with torch.no_grad():
            ms = timeit(lambda: warp.warp(a, c, c))
        print(json.dumps({"config": f"configs[1]: WarpModel GMM + TPS grid + grid_sample(border), batch {B}, fp16x3", "ms": round(ms, 4),
                          "frames_per_s": round(B * 1e3 / ms, 1), "algorithmic_gflop": round(9.333 * B, 1),
                          "tflops": round(9.333 * B / ms, 2), "frac_of_measured_bf16_peak": round(9.333 * B / ms / PEAK_TF, 4)}), flush=True)
    # configs[3]
    from shineon_virtual_tryon_b200.models.flownet import FlowNet

    torch.manual_seed(420)
    net = FlowNet()
    for m in net.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                torch.nn.init.uniform_(m.bias, -0.1, 0.1)
    net = net.to(dev).eval()
    B = 16
    im1, im2 = torch.rand(B, 3, 256, 192, generator=g).to(dev), torch.rand(B, 3, 256, 192, generator=g).to(dev)
    with torch.no_grad():
        ms_eager = timeit(lambda: net(im1, im2), iters=10)
        want = [t.clone() for t in net(im1, im2)]
        net.cuda_graph = True
        ms = timeit(lambda: net(im1, im2), iters=10)
        got = net(im1, im2)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(want, got)), "graph replay differs from the eager forward"
    print(json.dumps({"config": "configs[3]: FlowNet2 two-frame forward + confidence, batch 16 (xavier weights), fp16x3, CUDA-graph replay",
                      "ms": round(ms, 3), "ms_eager": round(ms_eager, 3),
                      "pairs_per_s": round(B * 1e3 / ms, 1), "algorithmic_gflop": round(49.56 * B, 1),
                      "tflops": round(49.56 * B / ms, 2), "frac_of_measured_bf16_peak": round(49.56 * B / ms / PEAK_TF, 4)}), flush=True)


def sams():
    """the reference's default SamsGenerator (64..1024 features, 3 middle blocks, agnostic + densepose + cloth
    label maps) on 256x192 frames; FLOPs counted from the module tree (2 * MAC of every conv the reference runs)."""
    import argparse

    from oracle import cases, weights
    from shineon_virtual_tryon_b200.networks.sams import SamsGenerator

    dev = torch.device("cuda", 0)
    hp = argparse.Namespace(**cases.SAMS_CASES["sams_default"][0])
    g = SamsGenerator(hp)
    shapes = {k: tuple(v.shape) for k, v in g.state_dict().items()}
    g.load_state_dict(weights.fix_spectral(weights.synth_state_dict(shapes, 420)), strict=True)
    g = g.to(dev).eval()
    gen = torch.Generator().manual_seed(1)
    for B in (1, 8):
        maps = {k: torch.randn(B, c, 256, 192, generator=gen).to(dev) for k, c in (("agnostic", 4), ("cloth", 3), ("densepose", 3))}
        from shineon_virtual_tryon_b200 import ops

        prof = []
        with torch.no_grad():
            ms = timeit(lambda: g(None, None, maps), iters=5)
            ops.PROFILE = prof
            g(None, None, maps)
            ops.PROFILE = None
            torch.cuda.synchronize()
        gf = sum(r[0] for r in prof) / 1e9
        tconv = sum(r[1].elapsed_time(r[2]) for r in prof)
        if "--layers" in sys.argv:
            agg = {}
            for r in prof:
                e = agg.setdefault(r[3], [0, 0.0, 0.0])
                e[0] += 1; e[1] += r[1].elapsed_time(r[2]); e[2] += r[0]
            for shp, (n, ms_, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                print(f"  {n:3d} x (N,H,W,Cin,cpad,Cout,k,s)={shp}: {ms_:7.3f} ms  {fl / ms_ / 1e9:7.1f} TF/s", file=sys.stderr)
        print(json.dumps({"config": f"SURVEY 8f N3: SamsGenerator default architecture, batch {B} x 256x192, fp16x3", "ms": round(ms, 3),
                          "frames_per_s": round(B * 1e3 / ms, 2), "conv_gflop": round(gf, 1), "conv_launches": len(prof),
                          "conv_ms_eager": round(tconv, 3), "tflops": round(gf / ms, 1),
                          "frac_of_measured_bf16_peak": round(gf / ms / PEAK_TF, 4)}), flush=True)


if __name__ == "__main__":
    if "--sams" in sys.argv:
        sams()
    else:
        main()
