"""In-memory two-stage try-on: GMM -> TPS warp -> TOM, frames batched along N.

The reference runs the two stages as separate `test.py` invocations and passes the warped cloth through 8-bit
PNGs on disk (docs/2_inference.md:13-39, models/warp_model.py:146-148, datasets/vvt_dataset.py:133-150).
This class chains the same two forwards (WarpModel.forward + grid_sample(border), UnetMaskModel.forward)
on the device; `run_host` is the host-buffer entry point used for the end-to-end number (pinned host
tensors in, pinned host tensor out, copies on the caller's stream).
"""
import torch


class TryOnPipeline:
    def __init__(self, warp_model, tom_model):
        self.warp_model = warp_model
        self.tom_model = tom_model
        self._host_out = None
        self._dev_in = None

    def set_precision(self, precision):
        self.warp_model.set_precision(precision)
        self.tom_model.set_precision(precision)

    @torch.no_grad()
    def __call__(self, person_gmm, cloth, person_tom):
        """person_gmm [F,22,H,W] (agnostic+cocopose), cloth [F,3,H,W], person_tom [F,7,H,W] (agnostic+densepose);
        all f32 CUDA.  Returns (p_tryon [F,3,H,W], tryon_mask [F,1,H,W], warped_cloth [F,3,H,W])."""
        warped_cloth, _, _, _ = self.warp_model.warp(person_gmm, cloth, cloth)
        _, tryon_masks, p_tryons, _ = self.tom_model(person_tom, warped_cloth)
        return p_tryons, tryon_masks, warped_cloth

    @torch.no_grad()
    def run_host(self, person_gmm_h, cloth_h, person_tom_h, out_h=None):
        """Host (pinned) tensors in, host (pinned) p_tryon out; H2D / D2H are part of the call."""
        dev = next(self.tom_model.parameters()).device
        shapes = (tuple(person_gmm_h.shape), tuple(cloth_h.shape), tuple(person_tom_h.shape))
        if self._dev_in is None or self._dev_in[0] != shapes:
            self._dev_in = (shapes, tuple(torch.empty(s, dtype=torch.float32, device=dev) for s in shapes))
        a, c, p = self._dev_in[1]
        a.copy_(person_gmm_h, non_blocking=True)
        c.copy_(cloth_h, non_blocking=True)
        p.copy_(person_tom_h, non_blocking=True)
        p_tryons, _, _ = self(a, c, p)
        if out_h is None:
            if self._host_out is None or self._host_out.shape != p_tryons.shape:
                self._host_out = torch.empty(p_tryons.shape, dtype=torch.float32).pin_memory()
            out_h = self._host_out
        out_h.copy_(p_tryons, non_blocking=True)
        return out_h
