"""Backward kernels on cuda:0 vs a plain PyTorch fp32 reference of the same op
(CPU autograd on the same seeded inputs)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import nchw, nhwc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from shineon_virtual_tryon_b200 import ops as _ops

    return _ops


def rel_err(got, want):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, f"shape {tuple(got.shape)} != {tuple(want.shape)}"
    assert torch.isfinite(got).all(), "non-finite values in the CUDA result"
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


# per-mode bound on max|err| / max|ref| for one GEMM-like gradient (operand rounding 2^-17 / 2^-9, f32 accumulate)
GRAD_TOL = {"bf16x3": 2e-4, "fp16x3": 2e-5, "bf16": 2e-2}

WGRAD_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad
    (2, 16, 12, 64, 64, 3, 1, 1),
    (3, 8, 6, 128, 192, 3, 1, 1),
    (2, 16, 12, 64, 128, 4, 2, 1),
    (5, 4, 3, 256, 64, 1, 1, 0),
    (1, 64, 48, 128, 4, 3, 1, 1),
    (2, 32, 24, 192, 100, 4, 2, 1),
    (7, 2, 2, 512, 512, 3, 1, 1),
]


def _ref_wgrad(x, g, w_shape, stride, pad):
    return torch.nn.grad.conv2d_weight(x, w_shape, g, stride=stride, padding=pad)


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "bf16"])
def test_conv2d_wgrad(ops, case, prec):
    N, H, W, Cin, Cout, k, s, p = case
    gen = torch.Generator().manual_seed(99 + Cin + Cout + k)
    x = torch.randn(N, Cin, H, W, generator=gen)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    g = torch.randn(N, Cout, Ho, Wo, generator=gen)
    if prec == "bf16":
        x, g = x.bfloat16().float(), g.bfloat16().float()
    want = _ref_wgrad(x, g, (Cout, Cin, k, k), s, p)
    xp = ops.nchw_to_planes(x.cuda(), prec=prec)
    gp = ops.nchw_to_planes(g.cuda(), prec=prec)
    gw = torch.full((Cout, Cin, k, k), float("nan"), device="cuda")
    ops.conv2d_wgrad(gp, xp, gw, Cout=Cout, Cin=Cin, kh=k, kw=k, stride=s, pad=p)
    assert rel_err(gw, want) < GRAD_TOL[prec]
    # accumulate form and a forced K-split
    ops.conv2d_wgrad(gp, xp, gw, Cout=Cout, Cin=Cin, kh=k, kw=k, stride=s, pad=p, alpha=0.5, beta=1.0, splits=3)
    assert rel_err(gw, 1.5 * want) < GRAD_TOL[prec]


def test_conv2d_wgrad_chan_map_and_bias(ops):
    """The up-conv's packed input is [skip pad64 | x' pad64]; the gradient must land on the true input channels."""
    N, H, W, c_skip, c_xp, Cout = 2, 8, 6, 40, 70, 96
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(N, c_skip + c_xp, H, W, generator=gen)
    g = torch.randn(N, Cout, H, W, generator=gen)
    want = _ref_wgrad(x, g, (Cout, c_skip + c_xp, 3, 3), 1, 1)
    p_skip, p_xp = ops.cpad64(c_skip), ops.cpad64(c_xp)
    buf = ops.Planes(N, H, W, p_skip + p_xp, prec="bf16x3", cpad=p_skip + p_xp)
    buf.hi.zero_()
    buf.lo.zero_()
    ops.nchw_to_planes(x[:, :c_skip].contiguous().cuda(), prec="bf16x3", out=buf.window(0, c_skip))
    ops.nchw_to_planes(x[:, c_skip:].contiguous().cuda(), prec="bf16x3", out=buf.window(p_skip, c_xp))
    cmap = [-1] * (p_skip + p_xp)
    for c in range(c_skip):
        cmap[c] = c
    for c in range(c_xp):
        cmap[p_skip + c] = c_skip + c
    gp = ops.nchw_to_planes(g.cuda(), prec="bf16x3")
    gw = torch.zeros(Cout, c_skip + c_xp, 3, 3, device="cuda")
    ops.conv2d_wgrad(gp, buf, gw, Cout=Cout, Cin=c_skip + c_xp, kh=3, kw=3, stride=1, pad=1, chan_map=cmap)
    assert rel_err(gw, want) < GRAD_TOL["bf16x3"]
    gb = torch.zeros(Cout, device="cuda")
    ops.channel_sum(nhwc(g).cuda(), gb)
    assert rel_err(gb, g.sum((0, 2, 3))) < 1e-5


def test_conv2d_wgrad_im2col_first_layer(ops):
    """Outermost down-conv (Cin 10, 4x4 s2): forward runs as a 1x1 GEMM over the im2col'd input; so does its wgrad."""
    N, H, W, Cin, Cout = 2, 32, 24, 10, 64
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(N, Cin, H, W, generator=gen)
    g = torch.randn(N, Cout, H // 2, W // 2, generator=gen)
    want = _ref_wgrad(x, g, (Cout, Cin, 4, 4), 2, 1)
    a = ops.nchw_im2col_planes(x.cuda(), None, 4, 4, 2, 1, prec="bf16x3")
    gp = ops.nchw_to_planes(g.cuda(), prec="bf16x3")
    gw = torch.zeros(Cout, Cin, 4, 4, device="cuda")
    ops.conv2d_wgrad(gp, a, gw, Cout=Cout, Cin=Cin, kh=4, kw=4, stride=2, pad=1, mode=1)
    assert rel_err(gw, want) < GRAD_TOL["bf16x3"]
