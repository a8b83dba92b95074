"""Diagnostics (GPU box): localise e2e mismatches layer by layer."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet as ounet  # noqa: E402
from shineon_virtual_tryon_b200 import ops  # noqa: E402
from shineon_virtual_tryon_b200.networks.attention.sagan import SelfAttention  # noqa: E402
from tests.test_kernels_gpu import CONV_CASES  # noqa: E402


def conv_cases():
    for split in ("fp16x3", "bf16x3", "fp16", "bf16"):
        for case in CONV_CASES[:6]:
            N, H, W, Cin, Cout, k, s, p = case
            g = torch.Generator().manual_seed(1234 + Cin + Cout + k)
            x = torch.randn(N, Cin, H, W, generator=g)
            w = torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)
            b = torch.randn(Cout, generator=g) * 0.1
            xp = ops.nchw_to_planes(x.cuda(), prec=split)
            pc = ops.PackedConv(w.cuda(), b.cuda(), stride=s, pad=p, prec=split)
            y, _ = ops.conv2d(xp, pc, want_f32=True)
            _, yp = ops.conv2d(xp, pc, want_planes=True)
            yd, _ = ops.conv2d(xp, pc, want_f32=True, direct=True)
            torch.cuda.synchronize()
            xs, ws = (x, w) if split.endswith('x3') else ((x.half().float(), w.half().float()) if split == 'fp16' else (x.bfloat16().float(), w.bfloat16().float()))
            want = F.conv2d(xs, ws, b, stride=s, padding=p)
            e1 = (y.permute(0, 3, 1, 2).cpu() - want).abs().max().item()
            e2 = (yd.permute(0, 3, 1, 2).cpu() - want).abs().max().item()
            e3 = (yp.float().cpu() - want).abs().max().item()
            print(f"conv {case} split={split}: igemm {e1:.2e} direct {e2:.2e} planes {e3:.2e} | max|want| {want.abs().max().item():.2f}")


def attention_cases():
    for (C, H, W) in [(64, 4, 3), (512, 4, 3), (512, 8, 6), (512, 16, 12), (864, 8, 6)]:
        g = torch.Generator().manual_seed(C + H)
        m = SelfAttention(C, "relu")
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.1))
            m.gamma.fill_(0.9)
        sd = {"a." + k: v.detach() for k, v in m.state_dict().items()}
        x = torch.randn(3, C, H, W, generator=g)
        want = ounet.self_attention(sd, "a.", x)
        got = m.cuda()(x.cuda())
        torch.cuda.synchronize()
        print(f"attention C={C} {H}x{W}: max err {(got.cpu() - want).abs().max().item():.2e} (max|want| {want.abs().max().item():.2f})")


def unet_levels():
    """Run the gelu+attention TOM case and compare every block's up-path output with the oracle."""
    from oracle import cases
    from tests.util import build_model
    import oracle.unet as ou

    model, sd = build_model("unet_mask")
    person, cloth, _ = cases.tom_inputs("tom_gelu_attn")
    x = torch.cat([person, cloth], 1)
    # oracle intermediates: monkeypatch _block to record outputs per level
    rec = {}
    orig = ou._block

    def spy(sd_, p, x_, level, *a, **k):
        out = orig(sd_, p, x_, level, *a, **k)
        rec[level] = out
        return out

    ou._block = spy
    with torch.no_grad():
        ou.unet_generator(sd, "unet.", x)
    ou._block = orig
    # ours: record Planes returned by each block.run
    from shineon_virtual_tryon_b200.networks.cpvton.unet import UnetSkipConnectionBlock

    mine = {}
    orig_run = UnetSkipConnectionBlock.run

    def spy_run(self, a_in, split):
        out = orig_run(self, a_in, split)
        lvl = 0
        b = model.unet.model
        while b is not self:
            b = b._parts["sub"]
            lvl += 1
        mine[lvl] = (a_in, out)
        return out

    UnetSkipConnectionBlock.run = spy_run
    with torch.no_grad():
        model(person.cuda(), cloth.cuda())
    torch.cuda.synchronize()
    UnetSkipConnectionBlock.run = orig_run
    for lvl in sorted(mine, reverse=True):
        a_in, out = mine[lvl]
        want = rec[lvl]
        if lvl == 0:
            got = out.permute(0, 3, 1, 2).cpu()
            print(f"level 0 final: max err {(got - want).abs().max().item():.2e}")
        else:
            # oracle block output = cat([x, x']) un-activated; ours = gelu(x') planes and a_in = gelu(x)
            cin = a_in.C
            xo, xp = want[:, :cin], want[:, cin:]
            e_in = (a_in.float().cpu() - F.gelu(xo)).abs().max().item()
            e_out = (out.float().cpu() - F.gelu(xp)).abs().max().item()
            print(f"level {lvl}: input(act) err {e_in:.2e}  x'(act) err {e_out:.2e}  |x'| max {xp.abs().max().item():.2f}")


if __name__ == "__main__":
    conv_cases()
    attention_cases()
    unet_levels()
