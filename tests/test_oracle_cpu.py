"""Pins the CPU oracle against the golden vectors produced by the reference's own classes
(oracle/make_golden.py), and checks the oracle's internal consistency.  CPU only."""
import pytest
import torch

from oracle import cases, flow_ops as fo, gmm, unet, weights
from tests.golden_util import load_golden
from tests.util import assert_close


def _tom_kwargs(over):
    return dict(n_frames=over.get("n_frames_total", 1), flow_warp=over.get("flow_warp", False), num_downs=6,
                num_attention=over.get("num_attn", 2), use_self_attn=over.get("self_attn", True),
                act=over.get("activation", "gelu"))


@pytest.mark.parametrize("name", list(cases.TOM_CASES))
def test_tom_oracle_matches_reference_golden(name):
    seed, shapes, gold = load_golden(name)
    sd = weights.synth_state_dict(shapes, seed)
    person, cloth, flows = cases.tom_inputs(name)
    with torch.no_grad():
        pr, tm, pt, fm = unet.tom_forward(sd, person, cloth, flows=flows, resample=fo.resample2d_fwd,
                                          **_tom_kwargs(cases.TOM_CASES[name][0]))
    # same ATen kernels, same graph: expect bit-level agreement (tolerance only for thread-count effects)
    st = cases.sub_step(name)
    assert_close(cases.subsample(pr, st), gold["p_rendereds"], atol=1e-5, rtol=1e-5, what="p_rendereds")
    assert_close(cases.subsample(tm, st), gold["tryon_masks"], atol=1e-5, rtol=1e-5, what="tryon_masks")
    assert_close(cases.subsample(pt, st), gold["p_tryons"], atol=1e-5, rtol=1e-5, what="p_tryons")
    if fm is not None:
        assert_close(cases.subsample(fm, st), gold["flow_masks"], atol=1e-5, rtol=1e-5, what="flow_masks")


@pytest.mark.parametrize("name", list(cases.GMM_CASES))
def test_gmm_oracle_matches_reference_golden(name):
    seed, shapes, gold = load_golden(name)
    sd = weights.synth_state_dict(shapes, seed)
    A, Bc, cloth, mask, theta = cases.gmm_inputs(name)
    t = gmm.TpsTables(256, 192, 5)
    with torch.no_grad():
        if theta is None:
            grid, theta = gmm.gmm_forward(sd, A, Bc, t)
        else:
            grid = gmm.tps_grid(theta, t)
    st = cases.sub_step(name)
    assert_close(theta, gold["theta"], atol=1e-5, rtol=1e-5, what="theta")
    assert_close(grid[:, ::st, ::st], gold["grid"], atol=1e-5, rtol=1e-5, what="grid")
    assert_close(cases.subsample(gmm.grid_sample(cloth, grid, "border"), st), gold["warped_cloth"], atol=1e-5, rtol=1e-5,
                 what="warped cloth")
    assert_close(cases.subsample(gmm.grid_sample(mask, grid, "zeros"), st), gold["warped_mask"], atol=1e-5, rtol=1e-5,
                 what="warped mask")


def test_pipeline_oracle_matches_reference_golden():
    """The benchmarked step's golden (80 frames through the reference's two models): the oracle chain reproduces the
    stored frames (every 5th) — run here on those frames only, every op on this path being per-sample."""
    from oracle import weights as W

    _, _, gold = load_golden("pipeline_b80")
    sdw = W.synth_state_dict(load_golden("gmm_b2")[1], 420)
    sdt = W.synth_state_dict(load_golden("tom_gelu_attn")[1], 420)
    pg, cloth, pt = cases.pipeline_inputs()
    fs, st = cases.PIPELINE_FRAME_STEP, cases.PIPELINE_SUB
    pg, cloth, pt = pg[::fs][:4], cloth[::fs][:4], pt[::fs][:4]  # 4 of the 16 stored frames keep the CPU suite short
    t = gmm.TpsTables(256, 192, 5)
    with torch.no_grad():
        grid, theta = gmm.gmm_forward(sdw, pg, cloth, t)
        wc = gmm.grid_sample(cloth, grid, "border")
        _, tm, ptry, _ = unet.tom_forward(sdt, pt, wc, **_tom_kwargs({}))
    # not bit-level like the other pins: the golden ran 16-frame chunks, this runs 4 frames, and ATen blocks its CPU convs
    # by batch size; the ~1e-5 difference in theta is then multiplied by the sampler (W/2 pixels x the cloth's gradient)
    assert_close(theta, gold["theta"][::fs][:4], atol=2e-5, rtol=1e-5, what="theta")
    assert_close(cases.subsample(wc, st), gold["warped_cloth"][:4], atol=3e-4, rtol=1e-4, what="warped cloth")
    assert_close(cases.subsample(tm, st), gold["tryon_masks"][:4], atol=3e-4, rtol=1e-4, what="tryon_masks")
    assert_close(cases.subsample(ptry, st), gold["p_tryons"][:4], atol=3e-4, rtol=1e-4, what="p_tryons")


def test_attention_levels_follow_reference_countdown():
    assert unet.attention_levels(6, 2, True) == {5, 4}
    assert unet.attention_levels(6, 3, True) == {5, 4, 3}
    assert unet.attention_levels(6, 2, False) == set()


# ---- the three CUDA-only ops: consistency of the restatement (backward == autograd of forward where the
# reference's backward is the exact gradient, i.e. away from its trunc-vs-floor quirk)
def test_correlation_backward_is_gradient_of_forward():
    g = torch.Generator().manual_seed(0)
    for (pad, k, maxd, s1, s2, C, H, W) in [(4, 1, 4, 1, 2, 4, 6, 5), (3, 3, 2, 1, 1, 3, 8, 7)]:
        a = torch.randn(1, C, H, W, generator=g, requires_grad=True)
        b = torch.randn(1, C, H, W, generator=g, requires_grad=True)
        o = fo.correlation_fwd(a, b, pad, k, maxd, s1, s2)
        assert o.shape[1:] == fo.correlation_out_shape(C, H, W, pad, k, maxd, s1, s2)
        go = torch.randn(o.shape, generator=g)
        o.backward(go)
        g1, g2 = fo.correlation_bwd(a.detach(), b.detach(), go, pad, k, maxd, s1, s2)
        assert_close(g1, a.grad, atol=1e-5, rtol=1e-4, what="d_in1")
        assert_close(g2, b.grad, atol=1e-5, rtol=1e-4, what="d_in2")


def test_flownetc_correlation_shape():
    # FlowNetC.py:31 config on the 256x192 path: 441 channels at 32x24
    assert fo.correlation_out_shape(256, 32, 24, 20, 1, 20, 1, 2) == (441, 32, 24)


def test_resample2d_properties():
    g = torch.Generator().manual_seed(1)
    img = torch.rand(2, 3, 12, 10, generator=g)
    zero = torch.zeros(2, 2, 12, 10)
    assert torch.equal(fo.resample2d_fwd(img, zero), img)  # identity flow
    shift = zero.clone()
    shift[:, 0] = 1.0  # sample one pixel to the right, clamped at the border
    want = torch.cat([img[..., 1:], img[..., -1:]], -1)
    assert torch.equal(fo.resample2d_fwd(img, shift), want)
    # backward == autograd where xf, yf >= 0 (no trunc/floor discrepancy, resample2d_kernel.cu:105-106)
    im = img.clone().requires_grad_()
    fl = (torch.rand(2, 2, 12, 10, generator=g) * 3).requires_grad_()
    o = fo.resample2d_fwd(im, fl)
    go = torch.randn(o.shape, generator=g)
    o.backward(go)
    g1, g2 = fo.resample2d_bwd(im.detach(), fl.detach(), go)
    assert_close(g1, im.grad, atol=1e-5, rtol=1e-4, what="d_img")
    assert_close(g2, fl.grad, atol=1e-5, rtol=1e-4, what="d_flow")


def test_channelnorm_backward_is_gradient():
    x = torch.randn(2, 3, 4, 5, requires_grad=True)
    o = fo.channelnorm_fwd(x)
    go = torch.randn_like(o)
    o.backward(go)
    assert_close(fo.channelnorm_bwd(x.detach(), o.detach(), go), x.grad, atol=1e-6, rtol=1e-5, what="d_in")


def test_flownet2_oracle_matches_reference_golden():
    from oracle import flownet2 as ofn

    seed, shapes, gold = load_golden("flownet2")
    sd = weights.synth_state_dict(shapes, seed)
    inp = cases.flownet2_inputs()
    with torch.no_grad():
        flow = ofn.flownet2(sd, inp)
        conf = ofn.flow_confidence(inp[:, :, 0], inp[:, :, 1], flow)
    assert_close(cases.subsample(flow, 2), gold["flow"], atol=1e-5, rtol=1e-5, what="flownet2 flow")
    assert torch.equal(cases.subsample(conf, 2), gold["conf"])


# ---- training step (row U6): the oracle's restatement of UnetMaskModel.training_step + autograd against the loss terms
# and parameter gradients the reference itself produced (oracle/make_golden.py train_golden)
def _train_kwargs(over):
    return dict(person_inputs=["agnostic", "densepose"], cloth_inputs=["cloth"], n_frames=over.get("n_frames_total", 1),
                flow_warp=over.get("flow_warp", False), num_downs=6, num_attention=over.get("num_attn", 2),
                use_self_attn=over.get("self_attn", True), act=over.get("activation", "gelu"))


@pytest.mark.parametrize("name", list(cases.TRAIN_CASES))
def test_training_oracle_matches_reference_golden(name):
    from oracle import train as otrain

    seed, shapes, gold = load_golden(name)
    sd = weights.synth_state_dict(shapes, seed)
    batch = cases.fold_frames(cases.train_batch(name))
    loss, comps, grads = otrain.tom_training_grads(sd, batch, **_train_kwargs(cases.TRAIN_CASES[name][0]))
    assert_close(loss.reshape(1), gold["loss"], atol=1e-5, rtol=1e-5, what="loss")
    for k in ("l1", "vgg", "tryon_mask_l1", "flow_mask_l1"):
        assert_close(comps[k].reshape(1), gold["log:loss/G/" + k], atol=1e-5, rtol=1e-5, what=k)
    keys = [k[6:] for k in gold if k.startswith("gnorm:")]
    assert sorted(keys) == sorted(grads), "the oracle must produce a gradient for exactly the parameters the reference trains"
    for k in keys:
        scale = gold["gnorm:" + k].item() / max(1.0, grads[k].numel() ** 0.5)  # rms of the reference gradient
        assert_close(cases.grad_sample(grads[k]), gold["gsamp:" + k], atol=1e-3 * scale + 1e-9, rtol=1e-3, what="grad " + k)


def test_model_state_dict_matches_reference_key_for_key():
    """Our UnetMaskModel mirrors the reference's module tree including the frozen VGG19 slices of criterionVGG."""
    from shineon_virtual_tryon_b200.models.unet_mask_model import UnetMaskModel
    from tests.util import make_hparams

    _, shapes, _ = load_golden("train_gelu_attn")
    mine = {k: tuple(v.shape) for k, v in UnetMaskModel(make_hparams(is_train=True)).state_dict().items()}
    assert mine == shapes


@pytest.mark.parametrize("shape", [(2, 5, 4, 3, 3), (1, 3, 1, 1, 2), (1, 4, 1, 5, 2), (1, 2, 6, 1, 3), (1, 3, 2, 2, 1)])
def test_lowres_decoder_formulation_equals_upsample_then_conv(shape):
    """The algebra behind ops.UpsampledConv3x3 / upconv3x3_gather: conv3x3(bilinear_up2(x)) from the
    tap-stacked products of the LOW-res tensor, borders (clamped interpolation, zero-padded conv) included."""
    import torch.nn.functional as F

    N, Cin, h, w, Cout = shape
    g = torch.Generator().manual_seed(h * 10 + w)
    x = torch.randn(N, Cin, h, w, generator=g)
    wt = torch.randn(Cout, Cin, 3, 3, generator=g)
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), wt, b, padding=1)
    got = unet.upsample_conv3x3_lowres(x, wt, b)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5), (got - want).abs().max()
