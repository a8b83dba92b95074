"""Oracle pin for the dataset-side frame prep: oracle/frame_prep.py vs the golden vectors produced by
the reference's own dataset methods (oracle/make_golden_prep.py) and vs Pillow / torchvision run in this process."""
import os

import numpy as np
import pytest

from oracle import frame_prep as fp

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frame_prep.npz"))
H, W = 256, 192


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_matches_reference_dataset_golden(seed):
    image, parse, cloth, densepose, pose = fp.synth_frame(seed, H, W)
    out = fp.frame_prep(image, parse, cloth, densepose, pose)
    g = lambda k: GOLD[f"s{seed}_{k}"]
    assert np.array_equal(out["image"][:, ::4, ::4], g("image_sub"))
    assert np.array_equal(out["cloth"][:, ::4, ::4], g("cloth_sub"))
    assert np.array_equal(out["cloth_mask"].astype(np.uint8), g("cloth_mask"))
    assert np.array_equal(out["im_head"][:, ::4, ::4], g("im_head_sub"))
    assert np.array_equal(out["silhouette"], g("silhouette"))          # bit-exact through two PIL resizes
    assert np.array_equal(out["agnostic"][0:1], g("silhouette"))
    assert tuple(g("pose_map_shape")) == out["cocopose"].shape
    assert out["cocopose"].min() == g("pose_map_minmax")[0] and out["cocopose"].max() == g("pose_map_minmax")[1] == -1.0
    assert np.array_equal(np.packbits(out["im_cocopose"][0] > 0), g("im_cocopose"))


def test_flo_decode_matches_reference_reader():
    assert np.array_equal(fp.decode_flo(GOLD["flo_bytes"].tobytes()), GOLD["flo_decoded"])
    bad = GOLD["flo_bytes"].copy()
    bad[0] ^= 1
    with pytest.raises(ValueError):
        fp.decode_flo(bad.tobytes())


@pytest.mark.parametrize("size", [(256, 192, 12, 16), (64, 48, 7, 5), (33, 17, 64, 40), (16, 12, 192, 256), (5, 9, 5, 30)])
def test_pil_bilinear_resize_restatement_is_bit_exact(size):
    """The third-party algorithm (Pillow Resample.c, 8-bit path): down- and up-scaling, odd sizes, one axis unchanged."""
    from PIL import Image

    h, w, ow, oh = size
    r = np.random.RandomState(h * 1000 + w)
    for img in (r.randint(0, 256, (h, w), dtype=np.uint8), (r.rand(h, w) > 0.5).astype(np.uint8) * 255):
        want = np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(fp.pil_resize_bilinear_u8(img, ow, oh), want)


def test_tensor_normalisation_matches_torchvision():
    import torch
    from PIL import Image
    from torchvision import transforms

    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    t = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])(Image.fromarray(img))
    assert torch.equal(t, torch.from_numpy(fp.norm_u8(img)))


def test_cocopose_squares_match_pil_draw():
    from PIL import Image, ImageDraw

    r = np.random.RandomState(3)
    for _ in range(20):
        pose = np.c_[r.uniform(-8, W + 8, 18), r.uniform(-8, H + 8, 18), r.rand(18)]
        im = Image.new("L", (W, H))
        d = ImageDraw.Draw(im)
        for x, y, _c in pose:
            if x > 1 and y > 1:
                d.rectangle((x - 5, y - 5, x + 5, y + 5), "white", "white")
        assert np.array_equal(fp.cocopose_vis_u8(pose, H, W, 5), np.array(im))


def test_library_host_coefficients_match_oracle():
    """shineon_pil_bilinear_coeffs is a pure host function of the C ABI (no GPU involved)."""
    import ctypes as C

    import torch

    from shineon_virtual_tryon_b200 import _lib

    lib = _lib.load()
    for a, b in ((192, 12), (256, 16), (12, 192), (16, 256), (33, 64), (17, 5), (7, 7)):
        ks = lib.shineon_pil_bilinear_coeffs(a, b, None, None)
        bo, kk = torch.zeros(b, 2, dtype=torch.int32), torch.zeros(b, ks, dtype=torch.int32)
        assert lib.shineon_pil_bilinear_coeffs(a, b, C.c_void_p(bo.data_ptr()), C.c_void_p(kk.data_ptr())) == ks
        ob, ok = fp.pil_resize_coeffs(a, b)
        assert np.array_equal(bo.numpy(), ob) and np.array_equal(kk.numpy(), ok)
    assert lib.shineon_pil_bilinear_coeffs(0, 4, None, None) < 0


def test_pil_resize_restatement_random_sizes():
    """Wider sweep of the Pillow restatement: random in/out sizes on both axes (down, up, identity, extreme ratios)."""
    from PIL import Image

    r = np.random.RandomState(99)
    for _ in range(40):
        h, w = int(r.randint(1, 70)), int(r.randint(1, 70))
        oh, ow = int(r.randint(1, 90)), int(r.randint(1, 90))
        img = r.randint(0, 256, (h, w), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        got = fp.pil_resize_bilinear_u8(img, ow, oh)
        assert np.array_equal(got, want), (h, w, oh, ow)


def test_silhouette_is_idempotent_on_uniform_maps_and_bounded():
    """Size-independent properties: an all-background map gives -1 everywhere, an all-foreground map +1, and the
    silhouette of any map stays inside [-1, 1]."""
    for h, w in ((256, 192), (64, 48), (128, 16)):
        assert np.all(fp.body_silhouette(np.zeros((h, w), np.uint8)) == -1.0)
        assert np.all(fp.body_silhouette(np.full((h, w), 7, np.uint8)) == 1.0)
        r = np.random.RandomState(h)
        s = fp.body_silhouette((r.rand(h, w) > 0.5).astype(np.uint8) * 3)
        assert s.min() >= -1.0 and s.max() <= 1.0 and s.shape == (1, h, w)
