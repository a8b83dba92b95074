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


# ------------------------------------------------------------------------------------------ pointwise backward
def _act_ref(name):
    return {None: lambda t: t, "relu": F.relu, "gelu": F.gelu, "swish": lambda t: t * torch.sigmoid(t),
            "sine": lambda t: torch.sin(30 * t), "leaky": lambda t: F.leaky_relu(t, 0.2)}[name]


@pytest.mark.parametrize("shape", [(2, 12, 8, 64), (3, 4, 3, 512), (1, 32, 24, 4), (2, 7, 5, 3)])
@pytest.mark.parametrize("act", [None, "gelu", "relu", "swish", "leaky"])
@pytest.mark.parametrize("do_norm", [True, False])
def test_instnorm_act_bwd(ops, shape, act, do_norm):
    N, H, W, C = shape
    gen = torch.Generator().manual_seed(C + H)
    x = (torch.randn(N, H, W, C, generator=gen) * 1.5 + 0.3).requires_grad_(True)
    g1 = torch.randn(N, H, W, C, generator=gen)
    g2 = torch.randn(N, H, W, C, generator=gen)
    xn = nchw(x)
    y = _act_ref(act)(F.instance_norm(xn, eps=1e-5) if do_norm else xn)
    # NB the upstream gradient is made contiguous in NCHW: torch's CPU instance_norm backward mishandles a
    # permuted (channels-last strided) grad_output when N == 1 (seen with torch 2.11)
    (y * nchw(g1 + g2)).sum().backward()
    xc = x.detach().cuda()
    ws = ops.instnorm_stats_ws(xc)
    ops.instnorm_act(xc, do_norm=do_norm, act=act, act_param=0.2, want_f32=True, want_planes=False, ws=ws)
    gx, gp = ops.instnorm_act_bwd(xc, ws if do_norm else None, g1.cuda(), g2.cuda(), do_norm=do_norm, act=act, act_param=0.2,
                                  want_f32=True, want_planes=True, prec="bf16x3")
    assert rel_err(gx, x.grad) < 2e-4
    assert rel_err(nhwc(gp.float()), x.grad) < 3e-4


def test_act_bwd(ops):
    gen = torch.Generator().manual_seed(3)
    z = torch.randn(3, 5, 7, 11, generator=gen).requires_grad_(True)
    g = torch.randn(3, 5, 7, 11, generator=gen)
    (F.gelu(z) * g).sum().backward()
    assert rel_err(ops.act_bwd(z.detach().cuda(), g.cuda(), None, act="gelu"), z.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 4, 3, 64, 64), (1, 16, 12, 10, 0), (2, 1, 1, 8, 3), (1, 2, 5, 7, 5)])
def test_upsample2x_cat_bwd(ops, shape):
    N, H, W, C0, C1 = shape
    gen = torch.Generator().manual_seed(H * W)
    s = torch.randn(N, C0 + C1, H, W, generator=gen).requires_grad_(True)
    g = torch.randn(N, 2 * H, 2 * W, C0 + C1, generator=gen)
    (F.interpolate(s, scale_factor=2, mode="bilinear", align_corners=False) * nchw(g)).sum().backward()
    g0, g1 = ops.upsample2x_cat_bwd(g.cuda(), C0, C1)
    want = nhwc(s.grad)
    assert rel_err(g0, want[..., :C0]) < 1e-5
    if C1:
        assert rel_err(g1, want[..., C0:]) < 1e-5


@pytest.mark.parametrize("shape", [(2, 4, 3, 64), (1, 16, 12, 512), (3, 8, 6, 128)])
def test_sagan_attention_bwd(ops, shape):
    N, H, W, C = shape
    Cq, HW = C // 8, H * W
    gen = torch.Generator().manual_seed(C)
    qkv = (torch.randn(N, H, W, 2 * Cq + C, generator=gen) * 0.5).requires_grad_(True)
    x = torch.randn(N, H, W, C, generator=gen)
    gamma = torch.tensor([0.8], requires_grad=True)
    g = torch.randn(N, H, W, C, generator=gen)
    t = qkv.view(N, HW, -1)
    q, k, v = t[..., :Cq], t[..., Cq:2 * Cq], t[..., 2 * Cq:]
    A = torch.softmax(q @ k.transpose(1, 2), -1)          # [N, i, j]
    out = gamma * (A @ v) + x.view(N, HW, C)
    (out * g.view(N, HW, C)).sum().backward()
    gg = torch.zeros(1, device="cuda")
    gq = ops.sagan_attention_bwd(qkv.detach().cuda(), gamma.detach().cuda(), g.cuda(), Cq, gg, beta_gamma=0.0)
    assert rel_err(gq, qkv.grad) < 1e-4
    assert rel_err(gg, gamma.grad) < 1e-4


@pytest.mark.parametrize("nf,flow_warp", [(1, False), (2, False), (2, True)])
def test_tom_compose_bwd(ops, nf, flow_warp):
    B, H, W = 2, 8, 6
    Cout = (5 if flow_warp else 4) * nf
    gen = torch.Generator().manual_seed(nf)
    u = torch.randn(B, H, W, Cout, generator=gen).requires_grad_(True)
    cloth = torch.rand(B, 3 * nf, H, W, generator=gen) * 2 - 1
    f = nf - 1
    wp = (torch.rand(B, 3, H, W, generator=gen) * 2 - 1).requires_grad_(True) if (flow_warp and f > 0) else None
    g_t = torch.randn(B, 3 * nf, H, W, generator=gen)
    g_m = torch.randn(B, nf, H, W, generator=gen)
    g_r = torch.randn(B, 3 * nf, H, W, generator=gen)
    g_f = torch.randn(B, nf, H, W, generator=gen) if flow_warp else None
    un = nchw(u)
    r = torch.tanh(un[:, 3 * f:3 * f + 3])
    m = torch.sigmoid(un[:, 3 * nf + f:3 * nf + f + 1])
    loss = (r * g_r[:, 3 * f:3 * f + 3]).sum() + (m * g_m[:, f:f + 1]).sum()
    rr = r
    if flow_warp:
        fm = torch.sigmoid(un[:, 4 * nf + f:4 * nf + f + 1])
        loss = loss + (fm * g_f[:, f:f + 1]).sum()
        if wp is not None:
            rr = (1 - fm) * wp + fm * r
    t = (1 - m) * rr + m * cloth[:, 3 * f:3 * f + 3]
    loss = loss + (t * g_t[:, 3 * f:3 * f + 3]).sum()
    loss.backward()
    gu = torch.zeros(B, H, W, Cout, device="cuda")
    gw = ops.tom_compose_bwd(u.detach().cuda(), cloth.cuda(), nf, flow_warp, gu, frame=f,
                             warped_prev=wp.detach().cuda() if wp is not None else None,
                             g_rendereds=g_r[:, 3 * f:3 * f + 3].contiguous().cuda(), g_masks=g_m[:, f:f + 1].contiguous().cuda(),
                             g_tryons=g_t[:, 3 * f:3 * f + 3].contiguous().cuda(),
                             g_flow_masks=g_f[:, f:f + 1].contiguous().cuda() if flow_warp else None,
                             want_g_warped=wp is not None)
    assert rel_err(gu, u.grad) < 1e-5  # channels of other frames stay zero in both
    if wp is not None:
        assert rel_err(gw, wp.grad) < 1e-5


def test_l1_loss_and_maxpool(ops):
    gen = torch.Generator().manual_seed(8)
    a = torch.randn(2, 3, 16, 12, generator=gen).requires_grad_(True)
    b = torch.randn(2, 3, 16, 12, generator=gen)
    (0.25 * F.l1_loss(a, b)).backward()
    loss = torch.full((1,), 2.0, device="cuda")
    ga = torch.zeros_like(a, device="cuda")
    ops.l1_loss(a.detach().cuda(), b.cuda(), loss, ga, weight=0.25, beta_loss=1.0)
    assert abs(loss.item() - 2.0 - 0.25 * F.l1_loss(a, b).item()) < 1e-5
    assert rel_err(ga, a.grad) < 1e-6
    x = torch.randn(2, 8, 6, 70, generator=gen).relu().requires_grad_(True)  # ties at zero, like post-ReLU VGG maps
    g = torch.randn(2, 4, 3, 70, generator=gen)
    y = F.max_pool2d(nchw(x), 2, 2)
    (y * nchw(g)).sum().backward()
    yf, yp = ops.maxpool2x2(x.detach().cuda(), prec="bf16x3")
    assert rel_err(yf, nhwc(y)) < 1e-6 and rel_err(nhwc(yp.float()), nhwc(y)) < 1e-4
    assert rel_err(ops.maxpool2x2_bwd(x.detach().cuda(), g.cuda()), x.grad) < 1e-6
