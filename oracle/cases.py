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
}
GMM_CASES = {"gmm_b2": dict(batch=2, theta_scale=None), "gmm_stress": dict(batch=2, theta_scale=0.3)}


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
