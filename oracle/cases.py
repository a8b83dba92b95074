"""Seeded parity cases shared by oracle/make_golden.py (reference side), the CPU tests (oracle side) and
the GPU tests (CUDA side).  Inputs depend only on the case name, so all three regenerate identical tensors."""
import torch

TOM_CASES = {
    # name: (hparams overrides, batch)
    "tom_gelu_attn": (dict(self_attn=True, num_attn=2, activation="gelu", ngf=64), 1),       # BASELINE config 1
    "tom_default_act": (dict(self_attn=False, num_attn=2, activation=None, ngf=64), 2),      # LeakyReLU-inplace quirk
    "tom_swish_attn3": (dict(self_attn=True, num_attn=3, activation="swish", ngf=64), 1),
    "tom_flow2": (dict(self_attn=True, num_attn=2, activation="gelu", n_frames_total=2, n_frames_now=2,
                       flow_warp=True), 1),                                                  # ngf=int(64*(ln2+1))=108
    # BASELINE config 3 secondary: channel-stacked 5-frame clip, 4 chained Resample2d blends, odd widths
    "tom_flow5": (dict(self_attn=True, num_attn=2, activation="gelu", n_frames_total=5, n_frames_now=5,
                       flow_warp=True), 1),                                                  # ngf=int(64*(ln5+1))=167
}
GMM_CASES = {"gmm_b2": dict(batch=2, theta_scale=None), "gmm_stress": dict(batch=2, theta_scale=0.3),
             "gmm_b8": dict(batch=8, theta_scale=None)}                                      # BASELINE configs[1]
# spatial subsampling step of the stored reference outputs (default 4)
SUBSAMPLE = {"gmm_b8": 8, "tom_flow5": 4}

# The benchmarked step itself (bench.py: 16 clips x 5 frames through TryOnPipeline): inputs for all 80 frames, reference
# outputs stored for every PIPELINE_FRAME_STEP-th frame at every PIPELINE_SUB-th pixel.
PIPELINE_FRAMES, PIPELINE_FRAME_STEP, PIPELINE_SUB = 80, 5, 8
FLOWNET2_B16 = 16  # BASELINE configs[3] batch


# SAMS generator: name -> (hparams overrides, batch, H, W)
_SAMS_BASE = dict(norm_G="spectralspadesyncbatch3x3", ngf_base=2, ngf_pow_outer=6, ngf_pow_inner=10, ngf_pow_step=1, num_middle=3,
                  attention_middle_indices=[], attention_decoder_indices=[], encoder_input="agnostic", n_frames_total=1, n_frames_now=1,
                  flow_warp=False, activation="gelu", person_inputs=["agnostic", "densepose"], cloth_inputs=["cloth"])
SAMS_CASES = {
    # three-level network, 3-frame window (two previous frames feed the encoder), spectral norm + eval-mode batch norm
    "sams_small": (dict(_SAMS_BASE, ngf_pow_outer=4, ngf_pow_inner=6, num_middle=2, n_frames_total=3, n_frames_now=3, flow_warp=True), 2, 64, 48),
    # instance norm, no spectral norm, LeakyReLU/ReLU activations, AttentiveMultiSpade in the middle (48 px) and the first
    # decoder layer (192 px), single-frame (zero previous frame), cocopose as a fourth label map
    "sams_instance_attn": (dict(_SAMS_BASE, norm_G="spadeinstance3x3", activation="relu", ngf_pow_outer=4, ngf_pow_inner=6,
                                num_middle=2, attention_middle_indices=["0"], attention_decoder_indices=["0"],
                                person_inputs=["agnostic", "densepose", "cocopose"], encoder_input="densepose"), 2, 32, 24),
    # the reference's default architecture (64 .. 1024 features, 3 middle blocks) on one 256x192 frame
    "sams_default": (dict(_SAMS_BASE), 1, 256, 192),
    # power step 2 that does not land on the inner / outer widths (the "extra layer" branches, sams_generator.py:158-165,
    # 197-209), 4 .. 128 features, 5x5 SPADE kernels on plain BatchNorm, Swish, a 2-frame window
    "sams_odd_widths": (dict(_SAMS_BASE, norm_G="spectralspadebatch5x5", activation="swish", ngf_pow_outer=3, ngf_pow_inner=6,
                             ngf_pow_step=2, num_middle=1, n_frames_total=2, n_frames_now=2, encoder_input="densepose"), 2, 32, 24),
}


def sams_inputs(name):
    """(prev_frames [b,n-1,3,h,w] | None, prev_labelmaps [b,n-1,c,h,w] | None, {input name: [b,c,h,w]})."""
    from oracle.sams import CHANNELS

    over, B, H, W = SAMS_CASES[name]
    g = _g(name)
    n = over["n_frames_total"]
    maps = {k: torch.randn(B, CHANNELS[k.upper()], H, W, generator=g) for k in sorted(over["person_inputs"] + over["cloth_inputs"])}
    if n == 1:
        return None, None, maps
    prev = torch.rand(B, n - 1, 3, H, W, generator=g) * 2 - 1
    prev_maps = torch.randn(B, n - 1, CHANNELS[over["encoder_input"].upper()], H, W, generator=g)
    return prev, prev_maps, maps


def sams_model_batch(name):
    """The batch dict SamsModel.generate_n_frames reads (models/sams_model.py:204-238): [b, n, c, h, w] per key."""
    from oracle.sams import CHANNELS

    over, B, H, W = SAMS_CASES[name]
    g = _g(name + ":model")
    n = over["n_frames_total"]
    keys = sorted(set(over["person_inputs"] + over["cloth_inputs"] + [over["encoder_input"]]))
    batch = {k: torch.randn(B, n, CHANNELS[k.upper()], H, W, generator=g) for k in keys}
    batch["image"] = torch.rand(B, n, 3, H, W, generator=g) * 2 - 1
    if over["flow_warp"]:
        batch["flow"] = torch.randn(B, n, 2, H, W, generator=g) * 2
    return batch


def _g(name):
    import zlib

    return torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2 ** 31))


def tom_inputs(name, H=256, W=192):
    over, B = TOM_CASES[name]
    n = over.get("n_frames_total", 1)
    g = _g(name)
    person = torch.randn(B, 7 * n, H, W, generator=g)
    cloth = torch.rand(B, 3 * n, H, W, generator=g) * 2 - 1
    flows = torch.randn(B, 2 * n, H, W, generator=g) * 3 if over.get("flow_warp") else None
    return person, cloth, flows


def gmm_inputs(name, H=256, W=192):
    cfg = GMM_CASES[name]
    g = _g(name)
    B = cfg["batch"]
    A = torch.randn(B, 22, H, W, generator=g)
    Bc = torch.randn(B, 3, H, W, generator=g)
    cloth = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.5).float()
    theta = None
    if cfg["theta_scale"] is not None:
        theta = (torch.rand(B, 50, generator=g) * 2 - 1) * cfg["theta_scale"]
    return A, Bc, cloth, mask, theta


def subsample(t, step=4):
    return t[..., ::step, ::step].contiguous()


def sub_step(name):
    return SUBSAMPLE.get(name, 4)


def pipeline_inputs(frames=PIPELINE_FRAMES, H=256, W=192):
    """(person_gmm [F,22,H,W], cloth [F,3,H,W], person_tom [F,7,H,W]) of the benchmarked step; agnostic (4 channels)
    is shared by both person tensors like in the reference's batches.  Clothes are smooth images in [-1,1]."""
    g = _g("pipeline_b80")
    agnostic = torch.randn(frames, 4, H, W, generator=g)
    cocopose = torch.randn(frames, 18, H, W, generator=g)
    densepose = torch.randn(frames, 3, H, W, generator=g)
    low = torch.rand(frames, 3, H // 8, W // 8, generator=g) * 2 - 1
    cloth = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
    cloth = (cloth + 0.05 * torch.randn(frames, 3, H, W, generator=g)).clamp(-1, 1)
    return torch.cat([agnostic, cocopose], 1), cloth, torch.cat([agnostic, densepose], 1)


def flownet2_inputs(B=1, H=256, W=192):
    """Two smooth-ish frames in [0,1] (the second a shifted / perturbed copy of the first) as [B,3,2,H,W]."""
    g = _g("flownet2")
    base = torch.nn.functional.interpolate(torch.rand(B, 3, H // 8 + 2, W // 8 + 2, generator=g), size=(H + 16, W + 16),
                                           mode="bilinear", align_corners=False)
    im1 = base[:, :, 8:8 + H, 8:8 + W] + 0.05 * torch.rand(B, 3, H, W, generator=g)
    im2 = base[:, :, 5:5 + H, 10:10 + W] + 0.05 * torch.rand(B, 3, H, W, generator=g)
    return torch.stack([im1, im2], dim=2).contiguous().clamp(0, 1)


# training cases: name -> (hparams overrides, batch size)
TRAIN_CASES = {
    "train_gelu_attn": (dict(self_attn=True, num_attn=2, activation="gelu", ngf=64, is_train=True), 2),
}


def train_batch(name, H=256, W=192):
    """The batch dict UnetMaskModel.training_step reads (unet_mask_model.py:137-146), [B, n, C, H, W] like the
    reference's collated dataset output; smooth-ish images so the L1 / perceptual terms are not pure noise."""
    over, B = TRAIN_CASES[name]
    n = over.get("n_frames_total", 1)
    g = _g(name)

    def smooth(c):
        low = torch.rand(B * n, c, H // 8, W // 8, generator=g) * 2 - 1
        t = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
        return (t + 0.1 * torch.randn(B * n, c, H, W, generator=g)).clamp(-1, 1).view(B, n, c, H, W)

    batch = dict(image=smooth(3), prev_image=smooth(3), cloth=smooth(3), agnostic=smooth(4), densepose=smooth(3),
                 cloth_mask=(smooth(1) > 0).float())
    if over.get("flow_warp"):
        batch["flow"] = torch.randn(B, n, 2, H, W, generator=g) * 3
    return batch


def fold_frames(batch):
    """[B, n, C, H, W] -> [B, n*C, H, W] (datasets/n_frames_interface.py:105-138)."""
    return {k: (v.reshape(v.shape[0], v.shape[1] * v.shape[2], *v.shape[3:]) if v.dim() == 5 else v) for k, v in batch.items()}


def grad_sample(t, n=256):
    """Fixed strided sample of a gradient tensor kept in the golden files."""
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].contiguous()
