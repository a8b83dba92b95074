"""Pins the oracle restatement of the three CUDA-only reference ops (and our kernels) against the REFERENCE'S OWN
kernels: oracle/_ref/*.so are the reference's unmodified .cu/.cc files compiled for sm_100 by oracle/build_ref.py.
Calls follow the reference's Python wrappers (resample2d.py:8-40, channelnorm.py:8-31, correlation.py:9-45)."""
import pytest
import torch

from oracle import build_ref, flow_ops as fo
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _ext(name):
    mod = build_ref.load(name)
    if mod is None:
        pytest.skip(f"oracle/_ref/{name}.so not built (reference sources only exist in the build container)")
    return mod


def test_reference_resample2d_kernel(cuda):
    ext = _ext("resample2d_cuda")
    from shineon_virtual_tryon_b200 import ops

    g = torch.Generator().manual_seed(21)
    B, C, H, W = 3, 3, 64, 48
    img = torch.rand(B, C, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g) * 4
    go = torch.randn(B, C, H, W, generator=g)
    for bilinear in (True, False):
        out = torch.zeros(B, C, H, W, device="cuda")
        ext.forward(img.cuda(), flow.cuda(), out, 1, bilinear)
        torch.cuda.synchronize()
        assert_close(fo.resample2d_fwd(img, flow, 1, bilinear), out, atol=1e-6, rtol=1e-5, what="oracle vs reference kernel (fwd)")
        assert_close(ops.resample2d_fwd(img.cuda(), flow.cuda(), 1, bilinear), out, atol=1e-6, rtol=1e-5,
                     what="ours vs reference kernel (fwd)")
    g1 = torch.zeros(B, C, H, W, device="cuda")
    g2 = torch.zeros(B, 2, H, W, device="cuda")
    ext.backward(img.cuda(), flow.cuda(), go.cuda(), g1, g2, 1, True)
    torch.cuda.synchronize()
    o1, o2 = fo.resample2d_bwd(img, flow, go)
    assert_close(o1, g1, atol=1e-5, rtol=1e-4, what="oracle vs reference kernel (d_in1)")
    assert_close(o2, g2, atol=1e-5, rtol=1e-4, what="oracle vs reference kernel (d_flow)")
    m1, m2 = ops.resample2d_bwd(img.cuda(), flow.cuda(), go.cuda())
    assert_close(m1, g1, atol=1e-5, rtol=1e-4, what="ours vs reference kernel (d_in1)")
    assert_close(m2, g2, atol=1e-5, rtol=1e-4, what="ours vs reference kernel (d_flow)")


def test_reference_channelnorm_kernel(cuda):
    ext = _ext("channelnorm_cuda")
    from shineon_virtual_tryon_b200 import ops

    g = torch.Generator().manual_seed(22)
    x = torch.randn(4, 3, 40, 28, generator=g)
    go = torch.randn(4, 1, 40, 28, generator=g)
    out = torch.zeros(4, 1, 40, 28, device="cuda")
    ext.forward(x.cuda(), out, 2)
    gi = torch.zeros(4, 3, 40, 28, device="cuda")
    ext.backward(x.cuda(), out, go.cuda(), gi, 2)
    torch.cuda.synchronize()
    assert_close(fo.channelnorm_fwd(x), out, atol=1e-6, rtol=1e-6, what="oracle vs reference kernel (fwd)")
    assert_close(fo.channelnorm_bwd(x, out.cpu(), go), gi, atol=1e-6, rtol=1e-5, what="oracle vs reference kernel (bwd)")
    mine = ops.channelnorm_fwd(x.cuda())
    assert_close(mine, out, atol=1e-6, rtol=1e-6, what="ours vs reference kernel (fwd)")
    assert_close(ops.channelnorm_bwd(x.cuda(), mine, go.cuda()), gi, atol=1e-6, rtol=1e-5, what="ours vs reference kernel (bwd)")


# pad_size >= max_displacement + kernel_radius in every config: the reference reads outside its padded buffers
# otherwise (correlation_cuda_kernel.cu:208, no bounds check); we define those taps as zero.
@pytest.mark.parametrize("cfg", [(20, 1, 20, 1, 2, 64, 32, 24), (5, 3, 4, 1, 2, 6, 10, 9), (4, 1, 4, 1, 1, 8, 12, 10)])
def test_reference_correlation_kernel(cuda, cfg):
    ext = _ext("correlation_cuda")
    from shineon_virtual_tryon_b200 import ops

    pad, k, maxd, s1, s2, C, H, W = cfg
    g = torch.Generator().manual_seed(23 + C)
    a = torch.randn(2, C, H, W, generator=g)
    b = torch.randn(2, C, H, W, generator=g)
    ac, bc = a.cuda(), b.cuda()
    rb1, rb2, out = ac.new(), bc.new(), ac.new()
    ext.forward(ac, bc, rb1, rb2, out, pad, k, maxd, s1, s2, 1)
    torch.cuda.synchronize()
    assert tuple(out.shape[1:]) == fo.correlation_out_shape(C, H, W, pad, k, maxd, s1, s2)
    assert_close(fo.correlation_fwd(a, b, pad, k, maxd, s1, s2), out, atol=1e-5, rtol=1e-4, what="oracle vs reference kernel (fwd)")
    assert_close(ops.correlation_fwd(ac, bc, pad, k, maxd, s1, s2), out, atol=1e-5, rtol=1e-4, what="ours vs reference kernel (fwd)")
    if H * W <= 130:
        go = torch.randn(out.shape, generator=g)
        rb1, rb2, g1, g2 = ac.new(), bc.new(), ac.new(), bc.new()
        ext.backward(ac, bc, rb1, rb2, go.cuda(), g1, g2, pad, k, maxd, s1, s2, 1)
        torch.cuda.synchronize()
        o1, o2 = fo.correlation_bwd(a, b, go, pad, k, maxd, s1, s2)
        assert_close(o1, g1, atol=1e-5, rtol=1e-4, what="oracle vs reference kernel (d_in1)")
        assert_close(o2, g2, atol=1e-5, rtol=1e-4, what="oracle vs reference kernel (d_in2)")
        m1, m2 = ops.correlation_bwd(ac, bc, go.cuda(), pad, k, maxd, s1, s2)
        assert_close(m1, g1, atol=1e-5, rtol=1e-4, what="ours vs reference kernel (d_in1)")
        assert_close(m2, g2, atol=1e-5, rtol=1e-4, what="ours vs reference kernel (d_in2)")
