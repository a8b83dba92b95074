"""Generates tests/golden/*.npz by running the UNMODIFIED reference classes (imported read-only through
oracle/ref_shim.py) on the seeded cases of oracle/cases.py.  Build-container only.

    python -m oracle.make_golden

Each fixture stores the state_dict key/shape list (weights are regenerated from (seed, key, shape) by
oracle/weights.py — `load_state_dict(strict=True)` into the reference module proves key compatibility),
and spatially subsampled reference outputs (every 4th pixel) to keep the files small.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, ref_shim, weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 420  # the reference's own seed (train.py:29)


def _save(name, shapes, **arrays):
    os.makedirs(OUT, exist_ok=True)
    meta = json.dumps({"seed": SEED, "shapes": {k: list(v) for k, v in shapes.items()}})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.frombuffer(meta.encode(), dtype=np.uint8),
                        **{k: v.detach().numpy() for k, v in arrays.items()})
    print("wrote", name, {k: tuple(v.shape) for k, v in arrays.items()})


def main():
    torch.manual_seed(SEED)
    ref_shim.install()
    ref_shim.patch_native_ops_with_oracle()  # Resample2d has no CPU path in the reference (tom_flow2 only)
    with torch.no_grad():
        for name, (over, _) in cases.TOM_CASES.items():
            m = ref_shim.build_unet_mask_model(**over)
            shapes = weights.shapes_of(m)
            m.load_state_dict(weights.synth_state_dict(shapes, SEED), strict=True)
            person, cloth, flows = cases.tom_inputs(name)
            pr, tm, pt, fm = m.forward(person, cloth, flows)
            st = cases.sub_step(name)
            arrs = dict(p_rendereds=cases.subsample(pr, st), tryon_masks=cases.subsample(tm, st), p_tryons=cases.subsample(pt, st))
            if fm is not None:
                arrs["flow_masks"] = cases.subsample(fm, st)
            _save(name, shapes, **arrs)
        import torch.nn.functional as F

        for name in cases.GMM_CASES:
            w = ref_shim.build_warp_model()
            shapes = weights.shapes_of(w)
            w.load_state_dict(weights.synth_state_dict(shapes, SEED), strict=True)
            A, Bc, cloth, mask, theta = cases.gmm_inputs(name)
            if theta is None:
                grid, theta_out = w.forward(A, Bc)
            else:  # stress the sampler with large offsets: gridGen + grid_sample only (warp_model.py:72,85-86)
                grid, theta_out = w.gridGen(theta), theta
            warped = F.grid_sample(cloth, grid, padding_mode="border")
            wmask = F.grid_sample(mask, grid, padding_mode="zeros")
            st = cases.sub_step(name)
            _save(name, shapes, theta=theta_out, grid=grid[:, ::st, ::st].contiguous(), warped_cloth=cases.subsample(warped, st),
                  warped_mask=cases.subsample(wmask, st))


def pipeline_golden():
    """The benchmarked step (bench.py): all 80 frames as ONE batch through the reference's WarpModel.forward ->
    F.grid_sample(border) -> UnetMaskModel.forward; outputs of every 5th frame at every 8th pixel are stored."""
    import torch.nn.functional as F

    ref_shim.install()
    with torch.no_grad():
        w = ref_shim.build_warp_model()
        wshapes = weights.shapes_of(w)
        w.load_state_dict(weights.synth_state_dict(wshapes, SEED), strict=True)
        m = ref_shim.build_unet_mask_model(**cases.TOM_CASES["tom_gelu_attn"][0])
        mshapes = weights.shapes_of(m)
        m.load_state_dict(weights.synth_state_dict(mshapes, SEED), strict=True)
        pg, cloth, pt = cases.pipeline_inputs()
        outs = []
        for s in range(0, pg.shape[0], 16):  # chunks only bound the CPU's memory; every op here is per-sample
            grid, theta = w.forward(pg[s:s + 16], cloth[s:s + 16])
            wc = F.grid_sample(cloth[s:s + 16], grid, padding_mode="border")
            pr, tm, ptry, _ = m.forward(pt[s:s + 16], wc)
            outs.append((theta, wc, tm, ptry))
        theta, wc, tm, ptry = (torch.cat(t) for t in zip(*outs))
        fs, st = cases.PIPELINE_FRAME_STEP, cases.PIPELINE_SUB
        _save("pipeline_b80", {}, theta=theta, warped_cloth=cases.subsample(wc[::fs], st),
              tryon_masks=cases.subsample(tm[::fs], st), p_tryons=cases.subsample(ptry[::fs], st))


def flownet2_golden(batch=1):
    """FlowNet2 (162.5 M parameters) through the reference module graph; its three CUDA-only ops run through
    oracle/flow_ops.py (which tests/test_ref_ext_gpu.py pins against the reference's own kernels).
    batch=16 is BASELINE configs[3]'s shape (stored at every 4th pixel, without the state_dict shape list)."""
    ref_shim.install()
    ref_shim.patch_native_ops_with_oracle()
    from models.flownet2_pytorch import models as fm
    from oracle import flownet2 as ofn

    with torch.no_grad():
        net = fm.FlowNet2().eval()
        shapes = weights.shapes_of(net)
        net.load_state_dict(weights.synth_state_dict(shapes, SEED), strict=True)
        inp = cases.flownet2_inputs(batch)
        flow = torch.cat([net(inp[i:i + 4]) for i in range(0, batch, 4)])  # per-sample network: chunks bound CPU memory
        conf = ofn.flow_confidence(inp[:, :, 0], inp[:, :, 1], flow)  # FlowNet.compute_flow_and_conf (flownet.py:55)
        if batch == 1:
            _save("flownet2", shapes, flow=cases.subsample(flow, 2), conf=cases.subsample(conf, 2))
        else:
            _save(f"flownet2_b{batch}", {}, flow=cases.subsample(flow, 4), conf=cases.subsample(conf, 4))


def train_golden():
    """Loss terms and parameter gradients of the reference's own UnetMaskModel.training_step + loss.backward()
    (unet_mask_model.py:137-217) with its real VGGLoss / Vgg19 classes (random weights instead of the ImageNet download)."""
    ref_shim.install()
    for name, (over, _) in cases.TRAIN_CASES.items():
        m = ref_shim.build_unet_mask_model(**over)
        m.train()
        shapes = weights.shapes_of(m)
        m.load_state_dict(weights.synth_state_dict(shapes, SEED), strict=True)
        m.visualize = lambda *a, **k: None
        res = m.training_step(cases.train_batch(name), 0)
        loss = res.minimize
        loss.backward()
        arrs = {"loss": loss.detach().reshape(1)}
        for k, v in res.items():
            arrs["log:" + k] = torch.as_tensor(v).detach().reshape(1)
        for k, p in m.named_parameters():
            if p.grad is not None:
                arrs["gnorm:" + k] = p.grad.norm().reshape(1)
                arrs["gsamp:" + k] = cases.grad_sample(p.grad)
        _save(name, shapes, **arrs)


if __name__ == "__main__":
    which = sys.argv[1:] or ["main", "flownet2", "flownet2_b16", "pipeline", "train"]
    if "main" in which:
        main()
    if "flownet2" in which:
        flownet2_golden()
    if "flownet2_b16" in which:
        flownet2_golden(cases.FLOWNET2_B16)
    if "pipeline" in which:
        pipeline_golden()
    if "train" in which:
        train_golden()
