"""CPU restatement of the SAMS generator forward (TEST INFRASTRUCTURE ONLY — nothing in the product path imports this).

Follows the reference module graph over torch.nn.functional, driven by a reference-keyed state_dict:
  SamsGenerator.forward            models/networks/sams/sams_generator.py:241-292
  make_encode_block / decode_block models/networks/sams/sams_generator.py:295-310
  AnySpadeResBlock.forward         models/networks/sams/spade.py:151-171
  SPADE.forward                    models/networks/sams/spade.py:68-84
  MultiSpade.forward               models/networks/sams/multispade.py:48-65
  AttentiveMultiSpade.forward      models/networks/sams/attentive_multispade.py:34-50
  SelfAttention.forward            models/networks/attention/sagan.py:29-53
  spectral_norm (eval)             torch.nn.utils.spectral_norm: W = weight_orig / (u . (W_mat v)), no power iteration
Pinned against the live reference by oracle/make_golden_sams.py -> tests/golden/sams_*.npz.
"""
import re

import torch
import torch.nn.functional as F

CHANNELS = dict(RGB=3, MASK=1, COCOPOSE=18, IM_HEAD=3, SILHOUETTE=1, AGNOSTIC=4, CLOTH=3, CLOTH_MASK=1, DENSEPOSE=3,
                FLOW=2)  # datasets/tryon_dataset.py:47-61


def _act(name, resblock):
    """spade.py:86-96 (SPADE: relu -> ReLU) and spade.py:173-183 (ResBlock: relu -> LeakyReLU(0.2))."""
    if name == "relu":
        return (lambda t: F.leaky_relu(t, 0.2)) if resblock else F.relu
    if name == "gelu":
        return F.gelu
    if name == "swish":
        return lambda t: t * torch.sigmoid(t)
    if name == "sine":
        return lambda t: torch.sin(30 * t)
    raise RuntimeError(name)


def _conv(sd, key, x, pad):
    """A conv that may be wrapped in spectral_norm (eval-mode weight, spade.py:138-143)."""
    if key + ".weight_orig" in sd:
        w0 = sd[key + ".weight_orig"]
        u, v = sd[key + ".weight_u"], sd[key + ".weight_v"]
        sigma = torch.dot(u, torch.mv(w0.reshape(w0.shape[0], -1), v))
        w = w0 / sigma
    else:
        w = sd[key + ".weight"]
    return F.conv2d(x, w, sd.get(key + ".bias"), padding=pad)


def _param_free_norm(sd, key, x, kind, train_stats=False):
    if kind == "instance":
        return F.instance_norm(x, eps=1e-5)
    if train_stats:
        return F.batch_norm(x, None, None, training=True, eps=1e-5)
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], training=False, eps=1e-5)


def spade(sd, key, x, seg, norm_kind, ks, activation):
    normalized = _param_free_norm(sd, key + ".param_free_norm", x, norm_kind)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = _act(activation, False)(F.conv2d(seg, sd[key + ".mlp_shared.0.weight"], sd[key + ".mlp_shared.0.bias"], padding=ks // 2))
    gamma = F.conv2d(actv, sd[key + ".mlp_gamma.weight"], sd[key + ".mlp_gamma.bias"], padding=ks // 2)
    beta = F.conv2d(actv, sd[key + ".mlp_beta.weight"], sd[key + ".mlp_beta.bias"], padding=ks // 2)
    return normalized * (1 + gamma) + beta


def any_spade(sd, key, x, seg, cfg):
    """SPADE / MultiSpade / AttentiveMultiSpade, told apart by the keys present (the reference picks the class per layer)."""
    norm_kind, ks, activation = cfg
    if key + ".mlp_shared.0.weight" in sd:  # plain SPADE (encoder)
        return spade(sd, key, x, seg, norm_kind, ks, activation)
    names = sorted({k[len(key) + len(".spade_layers."):].split(".")[0] for k in sd if k.startswith(key + ".spade_layers.")})
    if isinstance(seg, torch.Tensor):
        assert len(names) == 1
        seg = {names[0]: seg}
    assert len(seg) == len(names)
    if key + ".mlp_final.0.weight" in sd:  # AttentiveMultiSpade: parallel SPADEs, cat, attend, reduce
        outs = [spade(sd, f"{key}.spade_layers.{k}", x, s, norm_kind, ks, activation) for k, s in sorted(seg.items())]
        t = self_attention(sd, key + ".attention_layer", torch.cat(outs, 1))
        return F.leaky_relu(F.conv2d(t, sd[key + ".mlp_final.0.weight"], sd[key + ".mlp_final.0.bias"], padding=ks // 2), 0.01)
    for k, s in sorted(seg.items()):  # MultiSpade: sequential
        x = spade(sd, f"{key}.spade_layers.{k}", x, s, norm_kind, ks, activation)
    return x


def self_attention(sd, key, x):
    B, C, W, H = x.shape
    q = F.conv2d(x, sd[key + ".query_conv.weight"], sd[key + ".query_conv.bias"]).view(B, -1, W * H).permute(0, 2, 1)
    k = F.conv2d(x, sd[key + ".key_conv.weight"], sd[key + ".key_conv.bias"]).view(B, -1, W * H)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    v = F.conv2d(x, sd[key + ".value_conv.weight"], sd[key + ".value_conv.bias"]).view(B, -1, W * H)
    out = torch.bmm(v, att.permute(0, 2, 1)).view(B, C, W, H)
    return sd[key + ".gamma"] * out + x


def resblock(sd, key, x, seg, cfg):
    act = _act(cfg[2], True)
    learned = (key + ".conv_s.weight_orig" in sd) or (key + ".conv_s.weight" in sd)
    x_s = _conv(sd, key + ".conv_s", any_spade(sd, key + ".norm_s", x, seg, cfg), 0) if learned else x
    dx = _conv(sd, key + ".conv_0", act(any_spade(sd, key + ".spade_0", x, seg, cfg)), 1)
    dx = _conv(sd, key + ".conv_1", act(any_spade(sd, key + ".spade_1", dx, seg, cfg)), 1)
    return x_s + dx


def parse_norm_G(norm_G):
    m = re.search(r"spade(\D+)(\d)x\d", norm_G.replace("spectral", ""))
    kind = m.group(1)
    assert kind in ("instance", "syncbatch", "batch")
    return ("instance" if kind == "instance" else "batch"), int(m.group(2))


def generator_forward(sd, hp, prev_frames, prev_labelmaps, labelmaps):
    """sams_generator.py:241-292.  prev_*: [b, n, c, h, w] or None (n_frames_total == 1); labelmaps: {name: [b,c,h,w]}."""
    norm_kind, ks = parse_norm_G(hp.norm_G)
    cfg = (norm_kind, ks, hp.activation)
    if hp.n_frames_total > 1:
        b, n, c, h, w = prev_frames.shape
        x = prev_frames.reshape(b, -1, h, w)
        prev_labelmaps = prev_labelmaps.reshape(b, -1, h, w)
    else:
        ref = list(labelmaps.values())[0]
        b, _, h, w = ref.shape
        x = torch.zeros(b, 3 * max(hp.n_frames_total - 1, 1), h, w)
        prev_labelmaps = torch.zeros(b, CHANNELS[hp.encoder_input.upper()], h, w)

    def layer_kinds(prefix):
        idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")})
        return [(i, any(k.startswith(f"{prefix}.{i}.conv_0") for k in sd)) for i in idx]

    # encoder: Conv2d, then (ResBlock, Upsample(0.5)) pairs; nn.Upsample layers hold no state, so they are implied
    x = F.conv2d(x, sd["encode_layers.0.weight"], sd["encode_layers.0.bias"], padding=1)
    for i, is_block in layer_kinds("encode_layers")[1:]:
        x = resblock(sd, f"encode_layers.{i}", x, prev_labelmaps, cfg)
        x = F.interpolate(x, scale_factor=0.5, mode="nearest")
    for i, _ in layer_kinds("middle_layers"):
        x = resblock(sd, f"middle_layers.{i}", x, labelmaps, cfg)
    for i, is_block in layer_kinds("decode_layers"):
        if is_block:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = resblock(sd, f"decode_layers.{i}", x, labelmaps, cfg)
        else:
            x = F.conv2d(x, sd[f"decode_layers.{i}.weight"], sd[f"decode_layers.{i}.bias"], padding=1)
    return x


def prev_frames_and_maps(hp, batch, fIdx, all_G):
    """SamsModel.get_prev_frames_and_maps (models/sams_model.py:240-271): the window of previously generated frames (a ring
    over the frame buffer) and the encoder label maps of the frames before fIdx, zero-padded at the front."""
    enc = batch[hp.encoder_input]
    n = hp.n_frames_total
    if n == 1:
        return torch.zeros_like(all_G), torch.zeros_like(enc)
    n_prev = n - 1
    idx = torch.tensor([(i + 1) % n for i in range(fIdx, fIdx + n_prev)])
    prev = torch.index_select(all_G, 1, idx)
    b, _, c, h, w = enc.shape
    start = n_prev - fIdx
    maps = torch.cat((torch.zeros(b, start, c, h, w), enc[:, start:-1]), dim=1)
    return prev, maps


def generate_n_frames(sd, hp, batch, resample):
    """SamsModel.generate_n_frames (models/sams_model.py:204-238); frames before n_frames_total - n_frames_now stay zero
    (the reference's progressive-training window).  sd: generator.* keys
    stripped.  Returns (last frame, all generated frames [b,n,3,h,w])."""
    inputs = hp.person_inputs + hp.cloth_inputs
    all_G = torch.zeros_like(batch["image"])
    flows = torch.unbind(batch["flow"], dim=1) if hp.flow_warp else None
    fake = None
    for f in range(hp.n_frames_total - (getattr(hp, "n_frames_now", None) or hp.n_frames_total), hp.n_frames_total):
        maps = {k: batch[k][:, f] for k in inputs}
        prev, prev_maps = prev_frames_and_maps(hp, batch, f, all_G)
        out = generator_forward(sd, hp, prev, prev_maps, maps)
        fake, wmask = out[:, :3], out[:, 3:]
        if hp.flow_warp:
            last = all_G[:, f - 1].clone() if f > 0 else torch.zeros_like(all_G[:, f])
            warped = resample(last.contiguous(), flows[f].contiguous())
            fake = (1 - wmask) * warped + wmask * fake
        all_G[:, f] = fake
    return fake, all_G
