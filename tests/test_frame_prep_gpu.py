"""Dataset-side frame prep on the GPU vs the CPU oracle (pinned against the reference's dataset
methods and Pillow in tests/test_frame_prep_cpu.py) and vs the reference golden: byte / integer work, bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import frame_prep as fp

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frame_prep.npz"))
H, W = 256, 192


def _frames(seeds):
    fr = [fp.synth_frame(s, H, W) for s in seeds]
    st = lambda i, dt: torch.from_numpy(np.stack([f[i] for f in fr])).to(dt)
    return fr, st(0, torch.uint8), st(1, torch.uint8), st(2, torch.uint8), st(3, torch.uint8), st(4, torch.float64)


def test_frame_prep_bit_exact_vs_oracle_and_reference_golden(cuda):
    from shineon_virtual_tryon_b200 import ops

    seeds = [1, 2, 11, 12, 13]
    fr, image, parse, cloth, densepose, pose = _frames(seeds)
    prep = ops.FramePrep(H, W)
    out = prep(parse.cuda(), cloth.cuda(), densepose.cuda(), image.cuda(), pose=pose.cuda(), want_image=True)
    torch.cuda.synchronize()
    for i, (s, f) in enumerate(zip(seeds, fr)):
        want = fp.frame_prep(*f)
        for k in ("image", "cloth", "cloth_mask", "densepose", "agnostic", "cocopose", "im_cocopose"):
            assert np.array_equal(out[k][i].cpu().numpy(), want[k]), f"seed {s}: {k}"
        if s in (1, 2):  # the vectors produced by the reference's own dataset methods
            assert np.array_equal(out["agnostic"][i, 0:1].cpu().numpy(), GOLD[f"s{s}_silhouette"])
            assert np.array_equal(out["agnostic"][i, 1:, ::4, ::4].cpu().numpy(), GOLD[f"s{s}_im_head_sub"])
            assert np.array_equal(out["cloth"][i, :, ::4, ::4].cpu().numpy(), GOLD[f"s{s}_cloth_sub"])
            assert np.array_equal(np.packbits(out["im_cocopose"][i, 0].cpu().numpy() > 0), GOLD[f"s{s}_im_cocopose"])


def test_frame_prep_edge_cases(cuda):
    """All-background and all-foreground parse maps, extreme pixel values, no key points, other sizes."""
    from shineon_virtual_tryon_b200 import ops

    for h, w in ((256, 192), (64, 48), (32, 16)):
        prep = ops.FramePrep(h, w)
        r = np.random.RandomState(h)
        parse = np.stack([np.zeros((h, w), np.uint8), np.full((h, w), 13, np.uint8), r.randint(0, 20, (h, w)).astype(np.uint8)])
        img = np.stack([np.zeros((h, w, 3), np.uint8), np.full((h, w, 3), 255, np.uint8), r.randint(0, 256, (h, w, 3)).astype(np.uint8)])
        pose = np.zeros((3, 18, 3))
        t = lambda a: torch.from_numpy(a).cuda()
        out = prep(t(parse), t(img), t(img), t(img), pose=t(pose))
        for i in range(3):
            want = fp.frame_prep(img[i], parse[i], img[i], img[i], pose[i])
            for k in ("cloth", "cloth_mask", "densepose", "agnostic", "cocopose", "im_cocopose"):
                assert np.array_equal(out[k][i].cpu().numpy(), want[k]), f"{h}x{w} frame {i}: {k}"
    # a threshold inside [-1, 1] actually cuts (the default 240 never does, as in the reference)
    prep = ops.FramePrep(H, W, cloth_mask_threshold=0.5)
    _, image, parse, cloth, densepose, _ = _frames([3])
    m = prep(parse.cuda(), cloth.cuda(), densepose.cuda(), image.cuda())["cloth_mask"][0].cpu().numpy()
    want = fp.cloth_mask(fp.norm_u8(cloth[0].numpy()), 0.5)
    assert np.array_equal(m, want) and 0 < m.mean() < 1


def test_flo_decode(cuda):
    from shineon_virtual_tryon_b200 import ops

    got = ops.flo_decode(GOLD["flo_bytes"].tobytes())
    assert np.array_equal(got.cpu().numpy(), GOLD["flo_decoded"])
    bad = GOLD["flo_bytes"].copy()
    bad[1] ^= 4
    with pytest.raises(ValueError):
        ops.flo_decode(bad.tobytes())
    with pytest.raises(ValueError):
        ops.flo_decode(GOLD["flo_bytes"][:40].tobytes())


def test_raw_frame_entry_point_matches_tensor_entry_point(cuda):
    """TryOnPipeline.run_host_raw(uint8 frames) == run_host_batch(the reference-prepared f32 batch)."""
    from shineon_virtual_tryon_b200 import ops
    from shineon_virtual_tryon_b200.pipeline import TryOnPipeline
    from tests.util import build_model

    warp, _ = build_model("warp")
    tom, _ = build_model("unet_mask")
    pipe = TryOnPipeline(warp, tom)
    fr, image, parse, cloth, densepose, pose = _frames([21, 22])
    raw = {"image": image.pin_memory(), "parse": parse.pin_memory(), "cloth": cloth.pin_memory(), "densepose": densepose.pin_memory()}
    prep = ops.FramePrep(H, W)
    out, done = pipe.run_host_raw(raw, prep, u8_out=False)
    done.synchronize()
    got = out.clone()
    want_b = [fp.frame_prep(*f) for f in fr]
    batch = {k: torch.from_numpy(np.stack([b[k] for b in want_b])).pin_memory() for k in ("agnostic", "cocopose", "densepose", "cloth")}
    out2, done2 = pipe.run_host_batch(batch)
    done2.synchronize()
    assert torch.equal(got, out2)
    # 8-bit frames out (the default): exactly the reference writer's encoding of that f32 image, from the host entry
    # point and from the device-resident one, eager and replayed
    from oracle import image_io

    want_u8 = image_io.image_to_u8(got.numpy())
    out3, done3 = pipe.run_host_raw(raw, prep)
    done3.synchronize()
    assert out3.dtype == torch.uint8 and np.array_equal(out3.numpy(), want_u8)
    dev_raw = [raw[k].cuda() for k in pipe.RAW_KEYS]
    assert np.array_equal(pipe.run_raw(*dev_raw, prep).cpu().numpy(), want_u8)
    graphed = TryOnPipeline(warp, tom, cuda_graph=True)
    for _ in range(3):
        assert np.array_equal(graphed.run_raw(*dev_raw, prep).cpu().numpy(), want_u8)
        o, d = graphed.run_host_raw(raw, prep)
        d.synchronize()
        assert np.array_equal(o.numpy(), want_u8)
    assert graphed.replayed_launches > 0
    pipe.host_sync()
    graphed.host_sync()


def test_fused_prep_writes_the_stem_operands_bit_exactly(cuda):
    """ops.frame_prep_planes + ops.tps_warp_u8_planes == FramePrep -> torch.cat -> space-to-depth / im2col layout kernels
    (+ tps_grid_sample for the cloth), halfword for halfword, in both plane formats; and TryOnPipeline's fused raw path
    returns the same bytes as the unfused one, eager and replayed."""
    from shineon_virtual_tryon_b200 import ops
    from shineon_virtual_tryon_b200.pipeline import TryOnPipeline
    from tests.util import build_model

    fr, image, parse, cloth, densepose, pose = _frames([31, 32, 33])
    dev = [t.cuda() for t in (parse, cloth, densepose, image)]
    prep = ops.FramePrep(H, W)
    warp, _ = build_model("warp")
    tom, _ = build_model("unet_mask")
    for prec in ("fp16x3", "bf16x3", "bf16"):
        pr = ops.resolve_precision(prec)
        b = prep(*dev)
        fused = ops.frame_prep_planes(prep, *dev, prec=pr)
        # reference operands through the unfused kernels
        s2d_g = ops.S2dConv(torch.zeros(64, 22, 4, 4).cuda(), None, prec=pr)
        s2d_u = ops.S2dConv(torch.zeros(64, 10, 4, 4).cuda(), None, prec=pr)
        i2c = ops.Im2colConv(torch.zeros(64, 3, 4, 4).cuda(), None, 2, 1, prec=pr)
        want_g = s2d_g.prepare(torch.cat([b["agnostic"], b["cocopose"]], 1))
        want_c = i2c.prepare(b["cloth"])
        for got, want, what in ((fused["gmm_person"].planes, want_g, "gmm"), (fused["cloth_i2c"], want_c, "cloth im2col")):
            assert torch.equal(got.hi, want.hi), f"{prec} {what} hi"
            assert (got.lo is None and want.lo is None) or torch.equal(got.lo, want.lo), f"{prec} {what} lo"
        theta = (torch.rand(3, 50, generator=torch.Generator().manual_seed(3)) * 0.4 - 0.2).cuda()
        tables = warp.gridGen.tables(theta.device)
        warped = ops.tps_warp_u8_planes(theta, tables, dev[1], fused["unet_in"])
        (want_w,), _ = ops.tps_grid_sample(theta, tables, H, W, [(b["cloth"], "border")])
        assert torch.equal(warped, want_w), f"{prec} warped cloth"
        want_u = s2d_u.prepare(torch.cat([b["agnostic"], b["densepose"], want_w], 1))
        assert torch.equal(fused["unet_in"].planes.hi, want_u.hi), f"{prec} unet operand hi"
        assert want_u.lo is None or torch.equal(fused["unet_in"].planes.lo, want_u.lo), f"{prec} unet operand lo"
    # whole pipeline: fused == unfused
    pipe = TryOnPipeline(warp, tom)
    graphed = TryOnPipeline(warp, tom, cuda_graph=True)
    TryOnPipeline.FUSED_PREP = False
    try:
        want = pipe.run_raw(*dev, prep).clone()
    finally:
        TryOnPipeline.FUSED_PREP = True
    pipe2 = TryOnPipeline(warp, tom)
    assert torch.equal(pipe2.run_raw(*dev, prep), want)
    for _ in range(3):
        assert torch.equal(graphed.run_raw(*dev, prep), want)
