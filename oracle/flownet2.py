"""CPU oracle for FlowNet2 (reference: models/flownet2_pytorch/models.py:32-192, networks/FlowNetC.py:13-128,
FlowNetS.py:15-94, FlowNetSD.py:11-106, FlowNetFusion.py:11-67, submodules.py:7-38) and the FlowNet wrapper's
confidence map (models/flownet.py:42-63).  batchNorm=False, eval mode (only flow2 is returned by the sub-nets).
The three CUDA-only ops come from oracle/flow_ops.py."""
import torch
import torch.nn.functional as F

from . import flow_ops as fo


def _conv(sd, p, x, stride=1):
    """submodules.conv (batchNorm=False): Conv2d(pad=(k-1)//2, bias) + LeakyReLU(0.1)."""
    w = sd[p + ".0.weight"]
    return F.leaky_relu(F.conv2d(x, w, sd[p + ".0.bias"], stride=stride, padding=(w.shape[-1] - 1) // 2), 0.1)


def _iconv(sd, p, x):
    w = sd[p + ".0.weight"]
    return F.conv2d(x, w, sd.get(p + ".0.bias"), padding=(w.shape[-1] - 1) // 2)


def _deconv(sd, p, x):
    """submodules.deconv: ConvTranspose2d(4,2,1,bias) + LeakyReLU(0.1)."""
    return F.leaky_relu(F.conv_transpose2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=2, padding=1), 0.1)


def _pflow(sd, p, x):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)


def _upflow(sd, p, x):
    return F.conv_transpose2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=2, padding=1)


def _refine(sd, p, c6, c5, c4, c3, c2):
    """Shared decoder of FlowNetC / FlowNetS (FlowNetC.py:100-123, FlowNetS.py:68-89)."""
    flow6 = _pflow(sd, p + "predict_flow6", c6)
    cat5 = torch.cat((c5, _deconv(sd, p + "deconv5", c6), _upflow(sd, p + "upsampled_flow6_to_5", flow6)), 1)
    flow5 = _pflow(sd, p + "predict_flow5", cat5)
    cat4 = torch.cat((c4, _deconv(sd, p + "deconv4", cat5), _upflow(sd, p + "upsampled_flow5_to_4", flow5)), 1)
    flow4 = _pflow(sd, p + "predict_flow4", cat4)
    cat3 = torch.cat((c3, _deconv(sd, p + "deconv3", cat4), _upflow(sd, p + "upsampled_flow4_to_3", flow4)), 1)
    flow3 = _pflow(sd, p + "predict_flow3", cat3)
    cat2 = torch.cat((c2, _deconv(sd, p + "deconv2", cat3), _upflow(sd, p + "upsampled_flow3_to_2", flow3)), 1)
    return _pflow(sd, p + "predict_flow2", cat2)


def flownetc(sd, p, x):
    """FlowNetC.forward (FlowNetC.py:71-128)."""
    x1, x2 = x[:, 0:3], x[:, 3:]
    c1a = _conv(sd, p + "conv1", x1, 2)
    c2a = _conv(sd, p + "conv2", c1a, 2)
    c3a = _conv(sd, p + "conv3", c2a, 2)
    c3b = _conv(sd, p + "conv3", _conv(sd, p + "conv2", _conv(sd, p + "conv1", x2, 2), 2), 2)
    corr = F.leaky_relu(fo.correlation_fwd(c3a, c3b, 20, 1, 20, 1, 2), 0.1)  # FlowNetC.py:31,86-87
    redir = _conv(sd, p + "conv_redir", c3a)
    c31 = _conv(sd, p + "conv3_1", torch.cat((redir, corr), 1))
    c4 = _conv(sd, p + "conv4_1", _conv(sd, p + "conv4", c31, 2))
    c5 = _conv(sd, p + "conv5_1", _conv(sd, p + "conv5", c4, 2))
    c6 = _conv(sd, p + "conv6_1", _conv(sd, p + "conv6", c5, 2))
    return _refine(sd, p, c6, c5, c4, c31, c2a)


def flownets(sd, p, x):
    """FlowNetS.forward (FlowNetS.py:60-94)."""
    c1 = _conv(sd, p + "conv1", x, 2)
    c2 = _conv(sd, p + "conv2", c1, 2)
    c3 = _conv(sd, p + "conv3_1", _conv(sd, p + "conv3", c2, 2))
    c4 = _conv(sd, p + "conv4_1", _conv(sd, p + "conv4", c3, 2))
    c5 = _conv(sd, p + "conv5_1", _conv(sd, p + "conv5", c4, 2))
    c6 = _conv(sd, p + "conv6_1", _conv(sd, p + "conv6", c5, 2))
    return _refine(sd, p, c6, c5, c4, c3, c2)


def flownetsd(sd, p, x):
    """FlowNetSD.forward (FlowNetSD.py:66-106)."""
    c0 = _conv(sd, p + "conv0", x)
    c1 = _conv(sd, p + "conv1_1", _conv(sd, p + "conv1", c0, 2))
    c2 = _conv(sd, p + "conv2_1", _conv(sd, p + "conv2", c1, 2))
    c3 = _conv(sd, p + "conv3_1", _conv(sd, p + "conv3", c2, 2))
    c4 = _conv(sd, p + "conv4_1", _conv(sd, p + "conv4", c3, 2))
    c5 = _conv(sd, p + "conv5_1", _conv(sd, p + "conv5", c4, 2))
    c6 = _conv(sd, p + "conv6_1", _conv(sd, p + "conv6", c5, 2))
    flow6 = _pflow(sd, p + "predict_flow6", c6)
    cat5 = torch.cat((c5, _deconv(sd, p + "deconv5", c6), _upflow(sd, p + "upsampled_flow6_to_5", flow6)), 1)
    flow5 = _pflow(sd, p + "predict_flow5", _iconv(sd, p + "inter_conv5", cat5))
    cat4 = torch.cat((c4, _deconv(sd, p + "deconv4", cat5), _upflow(sd, p + "upsampled_flow5_to_4", flow5)), 1)
    flow4 = _pflow(sd, p + "predict_flow4", _iconv(sd, p + "inter_conv4", cat4))
    cat3 = torch.cat((c3, _deconv(sd, p + "deconv3", cat4), _upflow(sd, p + "upsampled_flow4_to_3", flow4)), 1)
    flow3 = _pflow(sd, p + "predict_flow3", _iconv(sd, p + "inter_conv3", cat3))
    cat2 = torch.cat((c2, _deconv(sd, p + "deconv2", cat3), _upflow(sd, p + "upsampled_flow3_to_2", flow3)), 1)
    return _pflow(sd, p + "predict_flow2", _iconv(sd, p + "inter_conv2", cat2))


def flownetfusion(sd, p, x):
    """FlowNetFusion.forward (FlowNetFusion.py:47-67)."""
    c0 = _conv(sd, p + "conv0", x)
    c1 = _conv(sd, p + "conv1_1", _conv(sd, p + "conv1", c0, 2))
    c2 = _conv(sd, p + "conv2_1", _conv(sd, p + "conv2", c1, 2))
    flow2 = _pflow(sd, p + "predict_flow2", c2)
    cat1 = torch.cat((c1, _deconv(sd, p + "deconv1", c2), _upflow(sd, p + "upsampled_flow2_to_1", flow2)), 1)
    flow1 = _pflow(sd, p + "predict_flow1", _iconv(sd, p + "inter_conv1", cat1))
    cat0 = torch.cat((c0, _deconv(sd, p + "deconv0", cat1), _upflow(sd, p + "upsampled_flow1_to_0", flow1)), 1)
    return _pflow(sd, p + "predict_flow0", _iconv(sd, p + "inter_conv0", cat0))


def flownet2(sd, inputs, prefix="", div_flow=20.0, rgb_max=1.0, stages=None):
    """FlowNet2.forward (models.py:127-192).  inputs [B,3,2,H,W] -> flow [B,2,H,W].
    `stages` (optional dict) receives intermediates for layer-wise checks."""
    p = prefix
    rgb_mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1).view(inputs.size()[:2] + (1, 1, 1))
    x = (inputs - rgb_mean) / rgb_max
    x = torch.cat((x[:, :, 0], x[:, :, 1]), dim=1)
    up = lambda t, mode: F.interpolate(t, scale_factor=4, mode=mode, **({"align_corners": False} if mode == "bilinear" else {}))

    def warp_block(flow):
        res = fo.resample2d_fwd(x[:, 3:].contiguous(), flow.contiguous())
        return res, fo.channelnorm_fwd(x[:, :3] - res)

    c_flow2 = flownetc(sd, p + "flownetc.", x)
    c_flow = up(c_flow2 * div_flow, "bilinear")
    res, nd = warp_block(c_flow)
    cat1 = torch.cat((x, res, c_flow / div_flow, nd), dim=1)
    s1_flow2 = flownets(sd, p + "flownets_1.", cat1)
    s1_flow = up(s1_flow2 * div_flow, "bilinear")
    res, nd = warp_block(s1_flow)
    cat2 = torch.cat((x, res, s1_flow / div_flow, nd), dim=1)
    s2_flow2 = flownets(sd, p + "flownets_2.", cat2)
    s2_flow = up(s2_flow2 * div_flow, "nearest")
    n_s2 = fo.channelnorm_fwd(s2_flow)
    _, d_s2 = warp_block(s2_flow)
    sd_flow2 = flownetsd(sd, p + "flownets_d.", x)
    sd_flow = up(sd_flow2 / div_flow, "nearest")
    n_sd = fo.channelnorm_fwd(sd_flow)
    _, d_sd = warp_block(sd_flow)
    cat3 = torch.cat((x[:, :3], sd_flow, s2_flow, n_sd, n_s2, d_sd, d_s2), dim=1)
    out = flownetfusion(sd, p + "flownetfusion.", cat3)
    if stages is not None:
        stages.update(x=x, c_flow2=c_flow2, s1_flow2=s1_flow2, s2_flow2=s2_flow2, sd_flow2=sd_flow2, cat1=cat1, cat3=cat3)
    return out


def flow_confidence(im1, im2, flow):
    """FlowNet.compute_flow_and_conf / norm (models/flownet.py:55,61-62): ||im1 - resample(im2, flow)||^2 < 0.02."""
    d = im1 - fo.resample2d_fwd(im2.contiguous(), flow.contiguous())
    return (torch.sum(d * d, dim=1, keepdim=True) < 0.02).float()


def compute_flow_and_conf(sd, im1, im2):
    """FlowNet.compute_flow_and_conf (models/flownet.py:42-59) including the bilinear pre/post resize taken when the
    HEIGHT is not a multiple of 64 (`if old_h != new_h`, :47,:56): nn.Upsample(size=..., mode='bilinear')."""
    old_h, old_w = im1.shape[2], im1.shape[3]
    new_h, new_w = old_h // 64 * 64, old_w // 64 * 64
    if old_h != new_h:
        im1 = F.interpolate(im1, size=(new_h, new_w), mode="bilinear", align_corners=False)
        im2 = F.interpolate(im2, size=(new_h, new_w), mode="bilinear", align_corners=False)
    flow = flownet2(sd, torch.cat([im1.unsqueeze(2), im2.unsqueeze(2)], dim=2))
    conf = flow_confidence(im1, im2, flow)
    if old_h != new_h:
        flow = F.interpolate(flow, size=(old_h, old_w), mode="bilinear", align_corners=False) * old_h / new_h
        conf = F.interpolate(conf, size=(old_h, old_w), mode="bilinear", align_corners=False)
    return flow, conf
