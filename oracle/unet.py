"""CPU oracle for the U-Net try-on module.  Reference: models/networks/cpvton/unet.py,
models/networks/attention/sagan.py, models/networks/activation.py, models/unet_mask_model.py."""
import torch
import torch.nn.functional as F


def activation(name, x, default):
    """_get_activation_fn (unet.py:201-211) / Sine, Swish (activation.py:4-18); `default` when --activation unset."""
    if name is None:
        return default(x)
    if name == "relu":
        return F.relu(x)
    if name == "gelu":
        return F.gelu(x)
    if name == "swish":
        return x * torch.sigmoid(x)
    if name == "sine":
        return torch.sin(30 * x)
    raise RuntimeError(name)


def self_attention(sd, p, x):
    """SelfAttention.forward (sagan.py:29-53)."""
    B, C, Wd, Ht = x.size()
    N = Wd * Ht
    q = F.conv2d(x, sd[p + "query_conv.weight"], sd[p + "query_conv.bias"]).view(B, -1, N).permute(0, 2, 1)
    k = F.conv2d(x, sd[p + "key_conv.weight"], sd[p + "key_conv.bias"]).view(B, -1, N)
    energy = torch.bmm(q, k)
    attn = F.softmax(energy, dim=-1)
    v = F.conv2d(x, sd[p + "value_conv.weight"], sd[p + "value_conv.bias"]).view(B, -1, N)
    out = torch.bmm(v, attn.permute(0, 2, 1)).view(B, C, Wd, Ht)
    return sd[p + "gamma"] * out + x


def _block(sd, p, x, level, num_downs, attn_levels, act, norm="instance"):
    """UnetSkipConnectionBlock.forward (unet.py:103-198).  level 0 = outermost, num_downs-1 = innermost.
    Sequential indices depend on which layers exist; they are derived here exactly like the constructor
    builds `model` (unet.py:137-184)."""
    outermost, innermost = level == 0, level == num_downs - 1
    has_attn = level in attn_levels

    def nrm(t, idx):
        if norm == "instance":
            return F.instance_norm(t, eps=1e-5)
        q = f"{p}{idx}."
        return F.batch_norm(t, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"], sd[q + "bias"],
                            False, 0.0, 1e-5)

    i = 0
    h = x
    if not outermost:
        # down_activation; LeakyReLU(0.2, inplace=True) mutates the skip tensor when no activation is named (unet.py:132)
        if act is None:
            x = F.leaky_relu(x, 0.2)
            h = x
        else:
            h = activation(act, x, None)
        i += 1
    bias = sd.get(f"{p}{i}.bias")
    h = F.conv2d(h, sd[f"{p}{i}.weight"], bias, stride=2, padding=1)
    i += 1
    if not outermost and not innermost:
        h = nrm(h, i)
        i += 1
    if has_attn:
        h = self_attention(sd, f"{p}{i}.", h)
        i += 1
    if not innermost:
        h = _block(sd, f"{p}{i}.model.", h, level + 1, num_downs, attn_levels, act, norm)
        i += 1
    h = activation(act, h, F.relu)  # up_activation (unet.py:134)
    i += 1
    h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=False)  # nn.Upsample (unet.py:138)
    i += 1
    h = F.conv2d(h, sd[f"{p}{i}.weight"], sd.get(f"{p}{i}.bias"), stride=1, padding=1)
    i += 1
    h = nrm(h, i)
    i += 1
    if has_attn:
        h = self_attention(sd, f"{p}{i}.", h)
        i += 1
    if outermost:
        return h
    return torch.cat([x, h], 1)


def attention_levels(num_downs, num_attention, use_self_attn):
    """Which blocks get SelfAttention: countdown from the innermost block (unet.py:38-94)."""
    if not use_self_attn:
        return set()
    return {num_downs - 1 - k for k in range(min(num_attention, num_downs)) }


def unet_generator(sd, prefix, x, num_downs=6, num_attention=2, use_self_attn=True, act="gelu", norm="instance"):
    """UnetGenerator.forward (unet.py:99-100)."""
    lv = attention_levels(num_downs, num_attention, use_self_attn)
    return _block(sd, prefix + "model.model.", x, 0, num_downs, lv, act, norm)


def tom_forward(sd, person, cloths, n_frames=1, flow_warp=False, flows=None, resample=None, **unet_kw):
    """UnetMaskModel.forward (unet_mask_model.py:64-135) -> (p_rendereds, tryon_masks, p_tryons, flow_masks)."""
    out = unet_generator(sd, "unet.", torch.cat([person, cloths], 1), **unet_kw)
    b3, b4 = 3 * n_frames, 4 * n_frames
    p_rendereds = torch.tanh(out[:, 0:b3])
    tryon_masks = torch.sigmoid(out[:, b3:b4])
    flow_masks = torch.sigmoid(out[:, b4:]) if flow_warp else None
    flows_c = list(torch.chunk(flows, n_frames, dim=1)) if flows is not None else None
    cl = list(torch.chunk(cloths, n_frames, dim=1))
    pr = list(torch.chunk(p_rendereds, n_frames, dim=1))
    tm = list(torch.chunk(tryon_masks, n_frames, dim=1))
    fm = list(torch.chunk(flow_masks, n_frames, dim=1)) if flow_masks is not None else None
    frames = []
    for f in range(n_frames):
        if flows_c is not None and f > 0:
            warped = resample(frames[f - 1], flows_c[f].contiguous())
            rend = (1 - fm[f]) * warped + fm[f] * pr[f]
        else:
            rend = pr[f]
        frames.append((1 - tm[f]) * rend + tm[f] * cl[f])
    return p_rendereds, tryon_masks, torch.cat(frames, dim=1), flow_masks


# ----------------------------------------------------------------------------- low-resolution form of upsample -> conv3x3
def _up3_coeffs(i, n):
    """The four coefficient triples over low-res neighbours (i-1, i, i+1) that give upsampled rows 2i-1, 2i, 2i+1, 2i+2
    (PyTorch bilinear x2, align_corners=False: clamped source index) with the conv's zero padding folded in."""
    first, last = i == 0, i == n - 1
    A = (0.0, 0.0, 0.0) if first else (0.75, 0.25, 0.0)
    B = (0.0, 1.0, 0.0) if first else (0.25, 0.75, 0.0)
    C = (0.0, 1.0, 0.0) if last else (0.0, 0.75, 0.25)
    D = (0.0, 0.0, 0.0) if last else (0.0, 0.25, 0.75)
    return A, B, C, D


def upsample_conv3x3_lowres(x, weight, bias=None):
    """CPU restatement of the product's decoder formulation (ops.UpsampledConv3x3 = tap-stacked 1x1 GEMM on the LOW-res
    tensor + csrc/upconv_gather.cu), used by tests to pin the algebra against the reference's own op order
    nn.Upsample(scale_factor=2, mode='bilinear') -> nn.Conv2d(3x3, padding=1) (models/networks/cpvton/unet.py:138-146):
        t[n, (fy,fx,co), i, j] = <x[n,:,i,j], w[co,:,fy,fx]>
        y[n, co, 2i+p, 2j+q]   = b[co] + sum_{fy,fx} sum_{a,b} cy[fy+p][a] * cx[fx+q][b] * t[n,(fy,fx,co), i-1+a, j-1+b]
    x: [N,Cin,h,w] -> [N,Cout,2h,2w]."""
    N, Cin, h, w = x.shape
    Cout = weight.shape[0]
    t = torch.einsum("ncij,ocyx->nyxoij", x.double(), weight.double())  # [N,3,3,Cout,h,w]
    tp = F.pad(t, (1, 1, 1, 1))  # neighbours outside the image meet zero coefficients
    y = torch.zeros(N, Cout, 2 * h, 2 * w, dtype=torch.float64)
    if bias is not None:
        y += bias.double().view(1, -1, 1, 1)
    for i in range(h):
        cy = _up3_coeffs(i, h)
        for j in range(w):
            cx = _up3_coeffs(j, w)
            nb = tp[:, :, :, :, i:i + 3, j:j + 3]  # [N,3,3,Cout,3(a),3(b)]
            for p in range(2):
                for q in range(2):
                    acc = 0
                    for fy in range(3):
                        wy = torch.tensor(cy[fy + p], dtype=torch.float64)
                        for fx in range(3):
                            wx = torch.tensor(cx[fx + q], dtype=torch.float64)
                            acc = acc + torch.einsum("noab,a,b->no", nb[:, fy, fx], wy, wx)
                    y[:, :, 2 * i + p, 2 * j + q] += acc
    return y.float()
