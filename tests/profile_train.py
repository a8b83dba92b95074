"""Timing of the native U-Net training step (GPU box).  Usage: python tests/profile_train.py [batch] [precision] [steps]
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list (tests/ncu_launch_table.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases  # noqa: E402  (test infrastructure: seeded batch only)
from shineon_virtual_tryon_b200.training import Trainer  # noqa: E402
from tests.util import build_model  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
    name = "train_gelu_attn"
    model, _ = build_model("unet_mask", **cases.TRAIN_CASES[name][0])
    model.train()
    model.set_train_precision(prec)
    tr = Trainer(model, lr=1e-4, cuda_graph=graph)
    b1 = cases.train_batch(name)
    reps = (B + 1) // 2
    batch = {k: torch.cat([v] * reps, 0)[:B].cuda() for k, v in b1.items()}
    for i in range(3):
        tr.train_batch(batch, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        res = tr.train_batch(batch, i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"train step B={B} {prec}{" cuda-graph" if graph else ""}: {ms:.3f} ms/step = {B / ms * 1e3:.1f} samples/s; loss {res['loss'].item():.4f}")


if __name__ == "__main__":
    main()
