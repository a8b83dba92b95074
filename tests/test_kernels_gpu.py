"""Per-kernel parity: every C-ABI entry point on cuda:0 vs the CPU oracle on the same seeded inputs."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close, nchw, nhwc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from shineon_virtual_tryon_b200 import ops as _ops

    return _ops


def _planes_from(ops, x_nchw, prec="fp16x3"):
    return ops.nchw_to_planes(x_nchw.cuda().contiguous(), prec=prec)


def _rounded(t, prec):
    """The operand rounding a single-product (non-split) mode applies."""
    return {"bf16": t.bfloat16().float(), "fp16": t.half().float()}.get(prec, t)


# per-mode bound of one conv layer against the fp32 oracle on O(1) data
CONV_TOL = {"fp16x3": dict(atol=3e-5, rtol=1e-4), "bf16x3": dict(atol=1e-4, rtol=1e-4),
            "fp16": dict(atol=5e-5, rtol=1e-4), "bf16": dict(atol=5e-5, rtol=1e-4)}


# ------------------------------------------------------------------------------------------ conv
CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad
    (1, 8, 16, 64, 64, 1, 1, 0),
    (2, 16, 12, 64, 128, 3, 1, 1),
    (3, 32, 24, 22, 64, 4, 2, 1),
    (2, 16, 12, 192, 512, 4, 2, 1),
    (5, 4, 3, 512, 512, 3, 1, 1),
    (2, 8, 6, 128, 640, 1, 1, 0),
    (1, 64, 48, 128, 4, 3, 1, 1),
    (2, 32, 32, 3, 64, 7, 2, 3),
    (2, 16, 16, 64, 128, 5, 2, 2),
    (1, 20, 12, 100, 70, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("prec", ["fp16x3", "bf16x3", "fp16", "bf16"])
def test_conv2d_igemm(ops, case, prec):
    N, H, W, Cin, Cout, k, s, p = case
    g = torch.Generator().manual_seed(1234 + Cin + Cout + k)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1
    xp = _planes_from(ops, x, prec)
    pc = ops.PackedConv(w.cuda(), b.cuda(), stride=s, pad=p, prec=prec)
    y, _ = ops.conv2d(xp, pc, want_f32=True)
    _, yp = ops.conv2d(xp, pc, want_planes=True)
    yd, _ = ops.conv2d(xp, pc, want_f32=True, direct=True)
    torch.cuda.synchronize()
    # split modes are compared with the exact fp32 conv; single-product modes with the conv of the rounded operands
    want = F.conv2d(_rounded(x, prec), _rounded(w, prec), b, stride=s, padding=p)
    tol = CONV_TOL[prec]
    assert_close(nchw(yd), want, what="direct (CUDA-core) vs oracle", **tol)
    assert_close(nchw(y), want, what="igemm (tcgen05) vs oracle", **tol)
    ptol = tol if prec.endswith("x3") else dict(atol=2e-2, rtol=1e-2)  # single 16-bit output plane
    assert_close(yp.float(), want, what="igemm planes vs oracle", **ptol)


@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, k, s, p: few output tiles, deep K (the 4x3 .. 8x6 levels) -> split-K
    (10, 8, 6, 512, 256, 4, 2, 1), (10, 4, 3, 256, 128, 3, 1, 1), (16, 4, 3, 1024, 1024, 3, 1, 1), (1, 16, 12, 512, 512, 3, 1, 1),
    (3, 8, 6, 256, 64, 3, 1, 1), (2, 4, 3, 2688, 72, 3, 1, 1)])
def test_conv2d_split_k(ops, case):
    """Split-K (K slices in workspace, last arriver finalises) against the single-pass kernel and the fp32 conv: f32 and
    planes outputs, fused InstanceNorm statistics, activation epilogue; repeated launches reuse the workspace (counters
    must come back to zero) and are bit-identical (fixed slice order)."""
    import ctypes as C

    from shineon_virtual_tryon_b200 import _lib

    N, H, W, Cin, Cout, k, s, p = case
    g = torch.Generator().manual_seed(77 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1
    xp = _planes_from(ops, x, "fp16x3")
    pc = ops.PackedConv(w.cuda(), b.cuda(), stride=s, pad=p, prec="fp16x3")
    want = F.conv2d(x, w, b, stride=s, padding=p)
    Ho, Wo = want.shape[2:]
    outs = {}
    for split in (False, True):
        ops.SPLIT_K = split
        try:
            ws = torch.zeros(2 * N * Cout, dtype=torch.float64, device="cuda")
            y, _ = ops.conv2d(xp, pc, want_f32=True, stats_ws=ws)
            _, yp = ops.conv2d(xp, pc, want_planes=True, post_act="gelu")
            y2, _ = ops.conv2d(xp, pc, want_f32=True)
        finally:
            ops.SPLIT_K = False
        torch.cuda.synchronize()
        outs[split] = (y, yp.float(), ws.clone(), y2)
    assert getattr(pc, "_sk_ws", None), "this shape was expected to take the split-K path"
    for split in (False, True):
        y, yp, ws, y2 = outs[split]
        assert_close(nchw(y), want, what=f"split_k={split} f32", **CONV_TOL["fp16x3"])
        assert_close(yp, F.gelu(want), what=f"split_k={split} planes + gelu", **CONV_TOL["fp16x3"])
        assert torch.equal(y, y2), "repeated launch on the same workspace differs"
        st = ws.view(N, Cout, 2).cpu()
        yc = nchw(y).double().cpu()
        assert_close(st[..., 0], yc.sum((2, 3)), atol=1e-4, rtol=1e-5, what="fused statistics: sum")
        assert_close(st[..., 1], (yc * yc).sum((2, 3)), atol=1e-4, rtol=1e-5, what="fused statistics: sum of squares")
    # slices are added in a fixed order: same accuracy class as the single pass (both within a few f32 ulps of each other)
    assert (outs[True][0] - outs[False][0]).abs().max().item() <= 2e-5 * want.abs().max().item()


def test_conv2d_epilogue_and_window(ops):
    """bias -> ReLU -> per-channel affine (folded BN) and writing into a channel window of a wider buffer."""
    g = torch.Generator().manual_seed(7)
    N, H, W, Cin, Cout = 2, 8, 6, 64, 64
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    sc = torch.rand(Cout, generator=g) + 0.5
    sh = torch.randn(Cout, generator=g) * 0.1
    xp = _planes_from(ops, x)
    pc = ops.PackedConv(w.cuda(), b.cuda(), stride=1, pad=1)
    out = ops.Planes(N, H, W, 192, device="cuda")
    out.hi.zero_(); out.lo.zero_()
    ops.conv2d(xp, pc, scale=sc.cuda(), shift=sh.cuda(), pre_act="relu", out_planes=out, out_coffset=64)
    torch.cuda.synchronize()
    want = F.relu(F.conv2d(x, w, b, padding=1)) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    got = out.float()
    assert_close(got[:, 64:128], want, atol=3e-5, rtol=1e-4, what="window")
    assert got[:, :64].abs().max().item() == 0 and got[:, 128:].abs().max().item() == 0


@pytest.mark.parametrize("cfg", [(2, 22, 64, 48, 64, 4, 2, 1), (2, 3, 32, 32, 64, 7, 2, 3), (1, 10, 64, 48, 64, 4, 2, 1),
                                 (2, 12, 64, 72, 64, 7, 2, 3), (3, 6, 40, 36, 64, 3, 1, 1), (1, 11, 33, 45, 64, 3, 1, 1)])
def test_im2col_first_layer(ops, cfg):
    """Small-Cin conv as im2col'd planes + dense 1x1 GEMM (cat of two NCHW inputs fused)."""
    N, Cin, H, W, Cout, k, s, p = cfg
    g = torch.Generator().manual_seed(Cin)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    c0 = Cin - Cin // 3
    conv = ops.Im2colConv(w.cuda(), b.cuda(), s, p)
    a = conv.prepare(x[:, :c0].contiguous().cuda(), x[:, c0:].contiguous().cuda() if c0 < Cin else None)
    y, _ = ops.conv2d(a, conv.pc, want_f32=True)
    want = F.conv2d(x, w, b, stride=s, padding=p)
    assert_close(nchw(y), want, atol=3e-5, rtol=1e-4, what="im2col conv (materialised planes)")
    # fused: im2col tile produced inside the kernel
    x0 = x[:, :c0].contiguous().cuda()
    x1 = x[:, c0:].contiguous().cuda() if c0 < Cin else None
    y2, p2 = conv.conv(x0, x1, want_f32=True, fused=True)
    assert_close(nchw(y2), want, atol=3e-5, rtol=1e-4, what="im2col conv (in-kernel producer)")
    _, p3 = conv.conv(x0, x1, post_act="leaky", act_param=0.1, want_planes=True, fused=True)
    assert_close(p3.float(), F.leaky_relu(want, 0.1), atol=3e-5, rtol=1e-4, what="im2col conv planes + leaky")


@pytest.mark.parametrize("cfg", [(2, 22, 64, 48, 64), (1, 10, 32, 24, 64), (3, 8, 8, 12, 96)])
@pytest.mark.parametrize("prec", ["fp16x3", "bf16"])
def test_s2d_first_layer(ops, cfg, prec):
    """4x4 s2 p1 small-Cin conv as a 2x2 s1 conv over shifted space-to-depth planes (cat of two NCHW inputs fused)."""
    N, Cin, H, W, Cout = cfg
    g = torch.Generator().manual_seed(Cin + H)
    x = _rounded(torch.randn(N, Cin, H, W, generator=g), prec)
    w = _rounded(torch.randn(Cout, Cin, 4, 4, generator=g) * 0.05, prec)
    b = torch.randn(Cout, generator=g) * 0.1
    c0 = Cin - Cin // 3
    conv = ops.first_layer_conv(w.cuda(), b.cuda(), 2, 1, prec=prec)
    assert isinstance(conv, ops.S2dConv)
    y, _ = conv.conv(x[:, :c0].contiguous().cuda(), x[:, c0:].contiguous().cuda(), pre_act="relu", want_f32=True)
    want = F.relu(F.conv2d(x, w, b, stride=2, padding=1))
    assert_close(nchw(y), want, what="s2d first-layer conv", **CONV_TOL[prec])
    # single input tensor, planes output
    _, yp = conv.conv(x.contiguous().cuda(), None, want_planes=True)
    assert_close(yp.float(), F.conv2d(x, w, b, stride=2, padding=1), atol=2e-2 if prec == "bf16" else 3e-5, rtol=1e-2 if prec == "bf16" else 1e-4,
                 what="s2d first-layer conv planes")


def test_tap_stacked_conv3x3(ops):
    g = torch.Generator().manual_seed(44)
    N, Cin, H, W, Cout = 2, 128, 32, 24, 4
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.03
    b = torch.randn(Cout, generator=g) * 0.1
    conv = ops.TapStackedConv3x3(w.cuda(), b.cuda())
    y = conv(_planes_from(ops, x))
    assert_close(nchw(y), F.conv2d(x, w, b, padding=1), atol=3e-5, rtol=1e-4, what="tap-stacked conv")


@pytest.mark.parametrize("shape", [(2, 128, 12, 9, 4), (1, 64, 1, 3, 5), (2, 192, 8, 6, 64), (1, 64, 5, 1, 36), (3, 64, 7, 11, 32),
                                   (1, 128, 1, 1, 96)])
def test_upsampled_conv3x3_at_low_resolution(ops, shape):
    """nn.Upsample(x2, bilinear) -> Conv2d(3x3, pad 1) (unet.py:138-146) from the tap-stacked low-res GEMM +
    upconv3x3_gather: must equal the conv over the materialised upsample, borders included (1-pixel-wide inputs too)."""
    g = torch.Generator().manual_seed(45)
    N, Cin, h, w, Cout = shape
    x = torch.randn(N, Cin, h, w, generator=g)
    wt = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.03
    b = torch.randn(Cout, generator=g) * 0.1
    want = F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), wt, b, padding=1)
    conv = ops.UpsampledConv3x3(wt.cuda(), b.cuda())
    y = conv(_planes_from(ops, x))
    assert tuple(y.shape) == (N, 2 * h, 2 * w, Cout)
    assert_close(nchw(y), want, atol=3e-5, rtol=1e-4, what="upsample->conv3x3 at low resolution")


@pytest.mark.parametrize("shape", [(2, 6, 8, 64, 32), (3, 4, 3, 130, 2), (16, 4, 3, 256, 200), (2, 24, 16, 64, 64),
                                   (5, 9, 7, 64, 144)])
def test_deconv_as_phase_convs(ops, shape):
    """ConvTranspose2d(4,2,1) (submodules.py:34-38) = 4 phase-wise 2x2 convolutions scattered into the output: as ONE launch
    (the phase is a tile dimension of conv_igemm, `deconv_phases`) and as four launches -- the two forms must agree bit for
    bit (same products, same accumulation order) and match torch within fp32-grade tolerance; f32 and plane outputs,
    Cout not a multiple of the channel tile (weight boxes run into the next phase's rows / past the tensor), several pixel
    tiles per phase, a channel window of a wider output buffer."""
    from shineon_virtual_tryon_b200.networks import deconv as dmod

    g = torch.Generator().manual_seed(11)
    N, H, W, Cin, Cout = shape
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cin, Cout, 4, 4, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    want = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    xp = _planes_from(ops, x)
    dc = dmod.PackedDeconv4x4s2(w.cuda(), b.cuda())
    outs = {}
    saved = dmod.MERGE_PHASES
    try:
        for merged in (True, False):
            dmod.MERGE_PHASES = merged
            y = dc(xp, want_f32=True)[0]
            cat = ops.Planes(N, 2 * H, 2 * W, 64 + ops.cpad64(Cout), device="cuda", cpad=64 + ops.cpad64(Cout))
            cat.hi.zero_()
            cat.lo.zero_()
            dc(xp, post_act="leaky", act_param=0.1, out_planes=cat.window(64, Cout))
            torch.cuda.synchronize()
            outs[merged] = (y.clone(), cat.hi.clone(), cat.lo.clone())
    finally:
        dmod.MERGE_PHASES = saved
    assert_close(nchw(outs[True][0]), want, atol=3e-5, rtol=1e-4, what="deconv (one launch)")
    for a, bb, what in zip(outs[True], outs[False], ("f32", "planes hi", "planes lo")):
        assert torch.equal(a, bb), f"deconv one launch vs four phase launches: {what} differ"
    got = (outs[True][1].float() + outs[True][2].float())[..., 64:64 + Cout].permute(0, 3, 1, 2).cpu()
    assert (outs[True][1][..., :64] == 0).all() and (outs[True][1][..., 64 + Cout:] == 0).all(), "wrote outside its window"
    assert_close(got, F.leaky_relu(want, 0.1), atol=3e-5, rtol=1e-4, what="deconv planes window")


@pytest.mark.parametrize("shape", [(2, 4, 3), (3, 16, 12), (1, 7, 5), (2, 64, 48)])
def test_flow_deconv_direct(ops, shape):
    """upsampled_flow* = ConvTranspose2d(2, 2, 4, 2, 1) of a predicted flow (FlowNetC.py:59-62) on the direct kernel: f32 NHWC
    flow in, the two channels of a concat window out (hi/lo planes); against torch, with and without bias, and nothing
    written outside the two channels."""
    g = torch.Generator().manual_seed(5)
    B, h, w = shape
    flow = torch.randn(B, 2, h, w, generator=g) * 3
    wt = torch.randn(2, 2, 4, 4, generator=g) * 0.3
    for bias in (None, torch.randn(2, generator=g)):
        want = F.conv_transpose2d(flow, wt, bias, stride=2, padding=1)
        cat = ops.Planes(B, 2 * h, 2 * w, 128, device="cuda", cpad=128)
        cat.hi.zero_(); cat.lo.zero_()
        ops.flow_deconv4x4s2_planes(flow.permute(0, 2, 3, 1).contiguous().cuda(), wt.cuda(), None if bias is None else bias.cuda(),
                                    cat.window(64, 2))
        torch.cuda.synchronize()
        got = (cat.hi.float() + cat.lo.float())[..., 64:66].permute(0, 3, 1, 2).cpu()
        assert_close(got, want, atol=3e-5, rtol=1e-4, what="flow deconv (direct kernel)")
        assert (cat.hi[..., :64] == 0).all() and (cat.hi[..., 66:] == 0).all() and (cat.lo[..., 66:] == 0).all()


# ------------------------------------------------------------------------------------------ gather ops
def _tps(ops, gs, H, W):
    from oracle.gmm import TpsTables

    t = TpsTables(H, W, gs)
    dev = ops.TpsTablesDev(t.Li, t.P_X, t.P_Y, t.grid_X[0, :], t.grid_Y[:, 0], gs, "cuda")
    return t, dev


@pytest.mark.parametrize("gs,scale", [(5, 0.1), (3, 0.3), (5, 0.6), (4, 0.2)])
def test_tps_grid_and_sample(ops, gs, scale):
    from oracle import gmm

    B, H, W = 3, 64, 48
    g = torch.Generator().manual_seed(gs)
    theta = (torch.rand(B, 2 * gs * gs, generator=g) * 2 - 1) * scale
    cloth = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.5).float()
    t, dev = _tps(ops, gs, H, W)
    want_grid = gmm.tps_grid(theta, t)
    got_grid = ops.tps_grid(theta.cuda(), dev, H, W)
    assert_close(got_grid, want_grid, atol=2e-5, rtol=1e-5, what="tps grid")
    # grid_sample on the ORACLE grid (isolates the sampler)
    for mode, src in (("border", cloth), ("zeros", mask)):
        want = gmm.grid_sample(src, want_grid, mode)
        got = ops.grid_sample(src.cuda(), want_grid.cuda(), mode)
        assert_close(got, want, atol=1e-5, rtol=1e-5, what=f"grid_sample {mode}")
    # fused path
    outs, grid2 = ops.tps_grid_sample(theta.cuda(), dev, H, W, [(cloth.cuda(), "border"), (mask.cuda(), "zeros")], want_grid=True)
    assert_close(grid2, want_grid, atol=2e-5, rtol=1e-5, what="fused grid")
    # fused = sampling at OUR grid: per-pixel-noise images amplify the ~1e-6 grid difference by (W/2 * |gradient|)
    assert_close(outs[0], gmm.grid_sample(cloth, want_grid, "border"), atol=1e-3, rtol=1e-2, what="fused cloth")
    assert_close(outs[1], gmm.grid_sample(mask, want_grid, "zeros"), atol=1e-3, rtol=1e-2, what="fused mask")


@pytest.mark.parametrize("bilinear", [True, False])
def test_resample2d(ops, bilinear):
    from oracle import flow_ops as fo

    g = torch.Generator().manual_seed(5)
    B, C, H, W = 3, 3, 40, 28
    img = torch.rand(B, C, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g) * 4
    want = fo.resample2d_fwd(img, flow, 1, bilinear)
    got = ops.resample2d_fwd(img.cuda(), flow.cuda(), 1, bilinear)
    assert_close(got, want, atol=1e-6, rtol=1e-5, what="resample2d fwd")
    if bilinear:
        go = torch.randn(B, C, H, W, generator=g)
        w1, w2 = fo.resample2d_bwd(img, flow, go)
        g1, g2 = ops.resample2d_bwd(img.cuda(), flow.cuda(), go.cuda())
        assert_close(g1, w1, atol=1e-5, rtol=1e-4, what="resample2d d_in1")
        assert_close(g2, w2, atol=1e-5, rtol=1e-4, what="resample2d d_flow")


@pytest.mark.parametrize("C,H,W", [(3, 256, 192), (1, 40, 68), (2, 33, 100), (4, 96, 64)])
def test_resample2d_halo_tile_kernel(ops, C, H, W):
    """The shared-memory halo-tile forward (csrc/resample_tile.cu: W >= 64, H >= 32, W % 4 == 0): ragged edge tiles, flows
    inside the 12-pixel window, far outside it (global-load fallback) and off the image (border clamp) -- bit-compatible
    accumulation order with the per-pixel kernel and the oracle."""
    from oracle import flow_ops as fo

    g = torch.Generator().manual_seed(C * 1000 + W)
    B = 2
    img = torch.rand(B, C, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g) * 4
    flow[:, :, ::7, ::5] *= 10            # far taps: outside the staged window, some outside the image
    flow[:, :, 3::11, 2::9] = 0.0         # integer positions (alpha = beta = 0)
    flow[0, 0, 0, 0], flow[0, 1, 0, 0] = -1e9, 1e9
    want = fo.resample2d_fwd(img, flow, 1, True)
    got = ops.resample2d_fwd(img.cuda(), flow.cuda(), 1, True)
    assert_close(got, want, atol=1e-6, rtol=1e-5, what="resample2d halo-tile fwd")


def test_channelnorm(ops):
    from oracle import flow_ops as fo

    g = torch.Generator().manual_seed(6)
    x = torch.randn(4, 3, 32, 24, generator=g)
    want = fo.channelnorm_fwd(x)
    got = ops.channelnorm_fwd(x.cuda())
    assert_close(got, want, atol=1e-6, rtol=1e-6, what="channelnorm fwd")
    go = torch.randn(4, 1, 32, 24, generator=g)
    assert_close(ops.channelnorm_bwd(x.cuda(), got, go.cuda()), fo.channelnorm_bwd(x, want, go), atol=1e-6, rtol=1e-5,
                 what="channelnorm bwd")


@pytest.mark.parametrize("cfg", [(20, 1, 20, 1, 2, 32, 16, 12), (5, 3, 4, 1, 2, 6, 10, 9), (4, 1, 4, 2, 1, 5, 12, 10)])
def test_correlation(ops, cfg):
    from oracle import flow_ops as fo

    pad, k, maxd, s1, s2, C, H, W = cfg
    g = torch.Generator().manual_seed(C)
    a = torch.randn(2, C, H, W, generator=g)
    b = torch.randn(2, C, H, W, generator=g)
    want = fo.correlation_fwd(a, b, pad, k, maxd, s1, s2)
    got = ops.correlation_fwd(a.cuda(), b.cuda(), pad, k, maxd, s1, s2, tensor_cores=False)
    assert_close(got, want, atol=1e-5, rtol=1e-4, what="correlation fwd (CUDA-core kernel)")
    if k == 1 and s1 == 1 and pad == maxd:
        got_tc = ops.correlation_fwd(a.cuda(), b.cuda(), pad, k, maxd, s1, s2, tensor_cores=True)
        assert_close(got_tc, want, atol=1e-5, rtol=1e-4, what="correlation fwd (tcgen05 GEMM + gather)")
        # the gather writing LeakyReLU(cost volume) as NHWC planes into a channel window (what FlowNetC consumes) must hold
        # exactly the planes the two-kernel route (NCHW f32 cost volume -> nchw_to_planes) produces, and nothing outside
        pa, pb = ops.nchw_to_planes(a.cuda()), ops.nchw_to_planes(b.cuda())
        DD = got_tc.shape[1]
        cat = ops.Planes(2, H, W, 64 + ops.cpad64(DD), device="cuda", cpad=64 + ops.cpad64(DD))
        cat.hi.zero_(); cat.lo.zero_()
        ops.correlation_planes(pa, pb, C, pad, maxd, s2, out_planes=cat.window(64, DD), act="leaky", act_param=0.1)
        ref = ops.nchw_to_planes(ops.correlation_planes(pa, pb, C, pad, maxd, s2), act="leaky", act_param=0.1)
        torch.cuda.synchronize()
        assert torch.equal(cat.hi[..., 64:64 + DD], ref.hi[..., :DD]) and torch.equal(cat.lo[..., 64:64 + DD], ref.lo[..., :DD])
        assert (cat.hi[..., :64] == 0).all() and (cat.hi[..., 64 + DD:] == 0).all() and (cat.lo[..., 64 + DD:] == 0).all()
    if H * W <= 120:
        go = torch.randn(want.shape, generator=g)
        w1, w2 = fo.correlation_bwd(a, b, go, pad, k, maxd, s1, s2)
        g1, g2 = ops.correlation_bwd(a.cuda(), b.cuda(), go.cuda(), pad, k, maxd, s1, s2)
        assert_close(g1, w1, atol=1e-5, rtol=1e-4, what="correlation d_in1")
        assert_close(g2, w2, atol=1e-5, rtol=1e-4, what="correlation d_in2")


# ------------------------------------------------------------------------------------------ norm / pointwise
@pytest.mark.parametrize("shape", [(2, 64, 32, 24), (3, 512, 4, 3), (2, 4, 64, 48), (2, 5, 16, 12), (1, 167, 8, 6)])
@pytest.mark.parametrize("act", [None, "gelu", "relu", "leaky", "swish", "sine"])
def test_instnorm_act(ops, shape, act):
    from oracle.unet import activation

    g = torch.Generator().manual_seed(shape[1])
    x = torch.randn(*shape, generator=g) * 2 + 0.7
    yf, yp = ops.instnorm_act(nhwc(x).cuda(), act=act, act_param=0.2, want_f32=True, want_planes=True)
    want = F.instance_norm(x, eps=1e-5)
    if act == "leaky":
        want = F.leaky_relu(want, 0.2)
    elif act is not None:
        want = activation(act, want, None)
    tol = dict(atol=3e-5, rtol=1e-4) if act != "sine" else dict(atol=2e-4, rtol=1e-3)
    assert_close(nchw(yf), want, what="instnorm f32", **tol)
    assert_close(yp.float(), want, what="instnorm planes", **tol)


def test_nchw_to_planes_and_upsample_cat(ops):
    g = torch.Generator().manual_seed(9)
    a = torch.randn(2, 70, 8, 6, generator=g)
    b = torch.randn(2, 30, 8, 6, generator=g)
    pa = ops.nchw_to_planes(a.cuda(), act="gelu")
    pb = ops.nchw_to_planes(b.cuda(), act="gelu")
    assert_close(pa.float(), F.gelu(a), atol=1e-5, rtol=1e-4, what="planes a")
    cat = ops.nchw_to_planes(a.cuda(), b.cuda())
    assert_close(cat.float(), torch.cat([a, b], 1), atol=1e-5, rtol=1e-4, what="planes cat")
    up = ops.upsample2x_cat(pa, pb)
    want = F.interpolate(torch.cat([F.gelu(a), F.pad(F.gelu(b), (0, 0, 0, 0, 0, 0))], 1), scale_factor=2, mode="bilinear",
                         align_corners=False)
    got = up.float()  # channels: [a (70) | pad to 128 | b (30) | pad]
    assert_close(got[:, :70], want[:, :70], atol=1e-5, rtol=1e-4, what="upsample a")
    assert_close(got[:, 128:158], want[:, 70:], atol=1e-5, rtol=1e-4, what="upsample b")
    assert got[:, 70:128].abs().max().item() == 0
    # extra ReLU on read (default-activation path)
    up2 = ops.upsample2x_cat(pa, None, act="relu")
    assert_close(up2.float()[:, :70], F.interpolate(F.relu(F.gelu(a)), scale_factor=2, mode="bilinear", align_corners=False),
                 atol=1e-5, rtol=1e-4, what="upsample relu")


@pytest.mark.parametrize("hw", [(4, 3), (8, 6), (16, 12)])
def test_sagan_attention(ops, hw):
    from oracle.unet import self_attention

    H, W = hw
    C, Cq, N = 64, 8, 3
    g = torch.Generator().manual_seed(H)
    x = torch.randn(N, C, H, W, generator=g)
    sd = {"a.query_conv.weight": torch.randn(Cq, C, 1, 1, generator=g) * 0.2, "a.query_conv.bias": torch.randn(Cq, generator=g) * 0.1,
          "a.key_conv.weight": torch.randn(Cq, C, 1, 1, generator=g) * 0.2, "a.key_conv.bias": torch.randn(Cq, generator=g) * 0.1,
          "a.value_conv.weight": torch.randn(C, C, 1, 1, generator=g) * 0.1, "a.value_conv.bias": torch.randn(C, generator=g) * 0.1,
          "a.gamma": torch.tensor([0.8])}
    want = self_attention(sd, "a.", x)
    q = F.conv2d(x, sd["a.query_conv.weight"], sd["a.query_conv.bias"])
    k = F.conv2d(x, sd["a.key_conv.weight"], sd["a.key_conv.bias"])
    v = F.conv2d(x, sd["a.value_conv.weight"], sd["a.value_conv.bias"])
    qkv = nhwc(torch.cat([q, k, v], 1)).cuda()
    yf, yp = ops.sagan_attention(qkv, nhwc(x).cuda(), sd["a.gamma"].cuda(), Cq, want_f32=True, want_planes=True)
    assert_close(nchw(yf), want, atol=2e-5, rtol=1e-4, what="attention f32")
    assert_close(yp.float(), want, atol=3e-5, rtol=1e-4, what="attention planes")


@pytest.mark.parametrize("hw", [(4, 3), (8, 6), (16, 12), (13, 10), (1, 1)])
@pytest.mark.parametrize("act", [None, "gelu"])
def test_sagan_attention_tensor_core(ops, hw, act):
    """The tcgen05 attention kernel (csrc/attention_tc.cu: C = 512, Cq = 64, HW <= 192 -- the U-Net's shapes): one and two
    query tiles, ragged tiles, key padding, energies with a wide dynamic range, f32 and plane outputs."""
    from oracle.unet import self_attention

    H, W = hw
    C, Cq, N = 512, 64, 3
    g = torch.Generator().manual_seed(H * 31 + W)
    x = torch.randn(N, C, H, W, generator=g)
    sd = {"a.query_conv.weight": torch.randn(Cq, C, 1, 1, generator=g) * 0.06, "a.query_conv.bias": torch.randn(Cq, generator=g) * 0.1,
          "a.key_conv.weight": torch.randn(Cq, C, 1, 1, generator=g) * 0.06, "a.key_conv.bias": torch.randn(Cq, generator=g) * 0.1,
          "a.value_conv.weight": torch.randn(C, C, 1, 1, generator=g) * 0.05, "a.value_conv.bias": torch.randn(C, generator=g) * 0.1,
          "a.gamma": torch.tensor([1.3])}
    want = self_attention(sd, "a.", x)
    if act == "gelu":
        want = F.gelu(want)
    q = F.conv2d(x, sd["a.query_conv.weight"], sd["a.query_conv.bias"])
    k = F.conv2d(x, sd["a.key_conv.weight"], sd["a.key_conv.bias"])
    v = F.conv2d(x, sd["a.value_conv.weight"], sd["a.value_conv.bias"])
    energy_span = torch.bmm(q.flatten(2).transpose(1, 2), k.flatten(2)).abs().max().item()
    qkv = nhwc(torch.cat([q, k, v], 1)).cuda()
    yf, yp = ops.sagan_attention(qkv, nhwc(x).cuda(), sd["a.gamma"].cuda(), Cq, act=act, want_f32=True, want_planes=True)
    torch.cuda.synchronize()
    err = assert_close(nchw(yf), want, atol=5e-5, rtol=1e-4, what="attention (tcgen05) f32")
    assert_close(yp.float(), want, atol=6e-5, rtol=1e-4, what="attention (tcgen05) planes")
    print(f"HW={H * W} act={act}: max abs err {err:.2e} (|energy| up to {energy_span:.1f})")


def test_l2norm_correlation_and_linear(ops):
    from oracle import gmm

    g = torch.Generator().manual_seed(3)
    B, C, h, w = 3, 512, 16, 12
    fa = torch.randn(B, C, h, w, generator=g).abs()
    fb = torch.randn(B, C, h, w, generator=g).abs()
    want = gmm.feature_correlation(gmm.feature_l2norm(fa), gmm.feature_l2norm(fb))  # [B, h*w, h, w]
    corr, planes = ops.l2norm_correlation(nhwc(fa).cuda(), nhwc(fb).cuda(), want_f32=True, want_planes=True)
    assert_close(nchw(corr), want, atol=2e-6, rtol=1e-4, what="l2norm+corr f32")
    assert_close(planes.float(), want, atol=1e-5, rtol=1e-4, what="l2norm+corr planes")
    # tensor-core form (per-image tcgen05 GEMM over the normalised planes), all plane formats; signed features too
    for fa2, fb2 in ((fa, fb), (torch.randn(B, C, h, w, generator=g), torch.randn(B, C, h, w, generator=g))):
        want2 = gmm.feature_correlation(gmm.feature_l2norm(fa2), gmm.feature_l2norm(fb2))
        for prec, tol in (("fp16x3", 2e-6), ("bf16x3", 2e-5), ("bf16", 2e-2)):
            c2, p2 = ops.l2norm_correlation_tc(nhwc(fa2).cuda(), nhwc(fb2).cuda(), prec=ops.resolve_precision(prec), want_f32=True)
            assert_close(nchw(c2), want2, atol=tol, rtol=1e-4, what=f"tensor-core l2norm+corr f32 ({prec})")
            assert_close(p2.float(), want2, atol=max(tol, 1e-5), rtol=1e-4, what=f"tensor-core l2norm+corr planes ({prec})")
    x = torch.randn(B, 64, 4, 3, generator=g)
    wt = torch.randn(50, 768, generator=g) * 0.05
    bs = torch.randn(50, generator=g) * 0.1
    want_t = torch.tanh(F.linear(x.view(B, -1), wt, bs))
    got_t = ops.linear_tanh(nhwc(x).cuda(), wt.cuda(), bs.cuda())
    assert_close(got_t, want_t, atol=1e-5, rtol=1e-4, what="linear+tanh")


@pytest.mark.parametrize("nf,flow", [(1, False), (2, True)])
def test_tom_compose(ops, nf, flow):
    g = torch.Generator().manual_seed(nf)
    B, H, W = 2, 16, 12
    cout = (5 if flow else 4) * nf
    u = torch.randn(B, cout, H, W, generator=g)
    cloth = torch.rand(B, 3 * nf, H, W, generator=g) * 2 - 1
    prev = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    pr = torch.empty(B, 3 * nf, H, W, device="cuda"); tm = torch.empty(B, nf, H, W, device="cuda")
    pt = torch.empty(B, 3 * nf, H, W, device="cuda"); fm = torch.empty(B, nf, H, W, device="cuda") if flow else None
    un = nhwc(u).cuda()
    for f in range(nf):
        ops.tom_compose(un, cloth.cuda(), nf, flow, (pr, tm, pt, fm), frame=f, warped_prev=prev.cuda() if (flow and f > 0) else None)
    w_pr = torch.tanh(u[:, :3 * nf]); w_tm = torch.sigmoid(u[:, 3 * nf:4 * nf])
    assert_close(pr, w_pr, atol=1e-6, rtol=1e-5, what="p_rendereds")
    assert_close(tm, w_tm, atol=1e-6, rtol=1e-5, what="tryon_masks")
    for f in range(nf):
        r = w_pr[:, 3 * f:3 * f + 3]
        if flow and f > 0:
            wf = torch.sigmoid(u[:, 4 * nf + f:4 * nf + f + 1])
            r = (1 - wf) * prev + wf * r
        want = (1 - w_tm[:, f:f + 1]) * r + w_tm[:, f:f + 1] * cloth[:, 3 * f:3 * f + 3]
        assert_close(pt[:, 3 * f:3 * f + 3], want, atol=2e-6, rtol=1e-5, what=f"p_tryon[{f}]")


def test_fused_adam_matches_torch_optim(ops):
    """shineon_adam_step vs torch.optim.Adam (CPU, fp32) over several steps, incl. the folded 1/world grad scale."""
    g = torch.Generator().manual_seed(77)
    n, world = 10007, 4
    p_ref = torch.nn.Parameter(torch.randn(n, generator=g))
    opt = torch.optim.Adam([p_ref], lr=1e-4)
    p = p_ref.detach().clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 6):
        grad_sum = torch.randn(n, generator=g) * world  # what the all-reduce (SUM) delivers
        p_ref.grad = grad_sum / world
        opt.step()
        ops.adam_step(p, grad_sum.cuda(), m, v, step, lr=1e-4, grad_scale=1.0 / world)
    assert_close(p, p_ref.detach(), atol=1e-6, rtol=1e-5, what="adam parameters after 5 steps")
