"""The 8-bit image writer (reference visualization.py:59-88).

CPU: the oracle (oracle/image_io.py) reproduces, bit for bit, PNG pixel arrays written by the reference's own
save_images (tests/golden/image_u8.npz).  GPU: shineon_image_to_u8 and the fused output of shineon_tom_compose are
bit-exact against the oracle on the same f32 input."""
import os

import numpy as np
import pytest
import torch

from oracle import image_io

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_u8.npz"))


@pytest.mark.parametrize("key", sorted(GOLD.files))
def test_oracle_matches_reference_pngs(key):
    seed, C = int(key[1:key.index("_")]), int(key[-1])
    got = image_io.image_to_u8(image_io.synth_images(seed, C=C))
    assert got.dtype == np.uint8 and np.array_equal(got, GOLD[key])


def test_oracle_edge_cases():
    x = np.array([-1.5, -1.0, -0.999999, 0.0, 0.999999, 1.0, 1.5], dtype=np.float32).reshape(1, 1, 1, 7)
    assert image_io.image_to_u8(x).reshape(-1).tolist() == [0, 0, 0, 127, 254, 255, 255]
    assert image_io.image_to_u8(np.zeros((0, 3, 4, 4), np.float32)).shape == (0, 4, 4, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 3, 32, 24), (2, 1, 32, 24), (80, 3, 256, 192), (1, 3, 5, 7), (2, 1, 3, 3)])
def test_image_to_u8_bit_exact(cuda, shape):
    from shineon_virtual_tryon_b200 import ops

    B, C, H, W = shape
    x = image_io.synth_images(11, B=B, C=C, H=H, W=W)
    got = ops.image_to_u8(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, image_io.image_to_u8(x))
    for key in GOLD.files:  # and the reference-written PNG pixels themselves
        seed, Cg = int(key[1:key.index("_")]), int(key[-1])
        xg = image_io.synth_images(seed, C=Cg)
        assert np.array_equal(ops.image_to_u8(torch.from_numpy(xg).cuda()).cpu().numpy(), GOLD[key])


@pytest.mark.gpu
@pytest.mark.parametrize("nf,flow,hw", [(1, False, (256, 192)), (2, True, (32, 24)), (1, False, (5, 7))])
def test_tom_compose_u8_output_is_the_encoded_f32_output(cuda, nf, flow, hw):
    """The fused 8-bit output of the compose kernel == image_to_u8 of the f32 p_tryon the same kernel writes."""
    from shineon_virtual_tryon_b200 import ops

    H, W = hw
    B = 3
    g = torch.Generator().manual_seed(5)
    un = (torch.randn(B, H, W, (5 if flow else 4) * nf, generator=g) * 2).cuda()
    cloth = (torch.rand(B, 3 * nf, H, W, generator=g) * 2.2 - 1.1).cuda()
    pr, tm, pt = torch.empty(B, 3 * nf, H, W).cuda(), torch.empty(B, nf, H, W).cuda(), torch.empty(B, 3 * nf, H, W).cuda()
    fm = torch.empty(B, nf, H, W).cuda() if flow else None
    u8 = torch.zeros(B, nf, H, W, 3, dtype=torch.uint8).cuda()
    for f in range(nf):
        prev = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).cuda() if (flow and f > 0) else None
        ops.tom_compose(un, cloth, nf, flow, (pr, tm, pt, fm), frame=f, warped_prev=prev, tryon_u8=u8)
        # u8-only call (no f32 outputs at all) writes the same bytes
        u8b = torch.zeros_like(u8)
        ops.tom_compose(un, cloth, nf, flow, (None, None, None, None), frame=f, warped_prev=prev, tryon_u8=u8b)
        assert torch.equal(u8b[:, f], u8[:, f])
    want = image_io.image_to_u8(pt.cpu().numpy().reshape(B * nf, 3, H, W)).reshape(B, nf, H, W, 3)
    assert np.array_equal(u8.cpu().numpy(), want)
