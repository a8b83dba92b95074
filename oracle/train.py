"""CPU oracle of the U-Net stage's training step (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates models/unet_mask_model.py:137-217 (UnetMaskModel.training_step: L1 + VGG perceptual + mask L1 + flow-mask
penalty on the last one/two frames), models/networks/loss.py:106-122 (VGGLoss) and models/networks/vgg.py:6-38 (Vgg19
slices of torchvision's vgg19.features) over torch.nn.functional; gradients come from torch autograd on the CPU.
Pinned by tests/golden/train_*.npz, produced by the reference's own training_step + loss.backward() (oracle/make_golden.py).
"""
import torch
import torch.nn.functional as F

from . import unet

# conv layers of torchvision vgg19.features[:30] by slice; "M" = MaxPool2d(2, 2)  (vgg.py:15-24)
VGG_SLICES = (
    ("slice1", ["0"]),
    ("slice2", ["2", "M", "5"]),
    ("slice3", ["7", "M", "10"]),
    ("slice4", ["12", "14", "16", "M", "19"]),
    ("slice5", ["21", "23", "25", "M", "28"]),
)
VGG_WEIGHTS = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]  # loss.py:112


def vgg19_features(sd, prefix, x):
    """Vgg19.forward (vgg.py:33-38) -> [h_relu1 .. h_relu5]."""
    outs, h = [], x
    for name, layers in VGG_SLICES:
        for l in layers:
            if l == "M":
                h = F.max_pool2d(h, 2, 2)
            else:
                h = F.relu(F.conv2d(h, sd[f"{prefix}{name}.{l}.weight"], sd[f"{prefix}{name}.{l}.bias"], padding=1))
        outs.append(h)
    return outs


def vgg_loss(sd, prefix, x, y):
    """VGGLoss.forward (loss.py:115-122)."""
    fx, fy = vgg19_features(sd, prefix, x), vgg19_features(sd, prefix, y)
    loss = 0
    for i in range(5):
        loss = loss + VGG_WEIGHTS[i] * F.l1_loss(fx[i], fy[i].detach())
    return loss


def tom_training_loss(sd, batch, person_inputs, cloth_inputs, n_frames=1, flow_warp=False, pen_flow_mask=1.0,
                      resample=None, **unet_kw):
    """training_step (unet_mask_model.py:137-191) -> (loss, {component: value}).  `batch` holds [B, n*C, H, W] tensors
    (already frame-folded, n_frames_interface.py:105-138)."""
    im, cm = batch["image"], batch["cloth_mask"]
    flow = batch["flow"] if flow_warp else None
    person = torch.cat([batch[k] for k in person_inputs], 1)
    cloths = torch.cat([batch[k] for k in cloth_inputs], 1)
    _, tryon_masks, p_tryons, flow_masks = unet.tom_forward(sd, person, cloths, n_frames=n_frames, flow_warp=flow_warp,
                                                            flows=flow, resample=resample, **unet_kw)
    pt = torch.chunk(p_tryons, n_frames, 1)
    tm = torch.chunk(tryon_masks, n_frames, 1)
    fm = torch.chunk(flow_masks, n_frames, 1) if flow_masks is not None else None
    im, cm = torch.chunk(im, n_frames, 1), torch.chunk(cm, n_frames, 1)
    two = n_frames > 1

    def pair(fn):
        cur = fn(-1)
        return 0.5 * (cur + fn(-2)) if two else cur

    l1 = pair(lambda i: F.l1_loss(pt[i], im[i]))
    vgg = pair(lambda i: vgg_loss(sd, "criterionVGG.vgg.", pt[i], im[i]))
    mask = pair(lambda i: F.l1_loss(tm[i], cm[i]))
    fml = (fm[-1].sum() if fm is not None else torch.zeros(())) * pen_flow_mask
    loss = l1 + vgg + mask + fml
    return loss, {"l1": l1, "vgg": vgg, "tryon_mask_l1": mask, "flow_mask_l1": fml}


def tom_training_grads(sd, batch, **kw):
    """-> (loss, components, {key: dLoss/dsd[key]}) for every U-Net parameter (the VGG slices are frozen, vgg.py:30-32)."""
    leaves = {k: v.clone().requires_grad_(k.startswith("unet.")) for k, v in sd.items()}
    loss, comps = tom_training_loss(leaves, batch, **kw)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return loss.detach(), {k: v.detach() for k, v in comps.items()}, grads
