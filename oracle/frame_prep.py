"""CPU restatement of the reference's per-frame dataset tensor prep.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product path.

Follows datasets/tryon_dataset.py of the reference (file:line cited per function).  The image resampling the reference
does through PIL (`Image.resize(..., Image.BILINEAR)`, tryon_dataset.py:352-358) lives in a third-party dependency
that is not part of /root/reference: Pillow (the reference pins pillow=7.2.0, sams-pt1.6.yaml; this container has
12.2.0 — the 8-bit resampling algorithm, src/libImaging/Resample.c, is unchanged between them).  Its published
algorithm is restated in `pil_resize_coeffs` / `pil_resize_bilinear_u8`.

Parity pin: tests/test_oracle_cpu.py checks every function here against (a) Pillow itself and torchvision's
ToTensor/Normalize in this container and (b) the golden vectors in tests/golden/frame_prep.npz, which
oracle/make_golden_prep.py produced by calling the UNMODIFIED reference methods (TryonDataset.get_person_head,
get_person_body_silhouette, get_input_cloth_mask, convert_pose_data_to_pose_map_and_vis, readFlow) on seeded inputs.

Integer / byte arithmetic throughout: the bar is bit-exact.
"""
import math

import numpy as np

# LIP labels (tryon_dataset.py:21-41) that get_person_head keeps (tryon_dataset.py:326-341)
LIP_HEAD_LABELS = (1, 2, 4, 13, 8, 9, 11, 12, 16, 17, 18, 19)
PRECISION_BITS = 32 - 8 - 2  # Pillow Resample.c


def norm_u8(img_hwc):
    """transforms.ToTensor() + Normalize(0.5, 0.5) (tryon_dataset.py:109-119,149-152): uint8 HWC (or HW) -> f32 CHW."""
    a = np.asarray(img_hwc, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    t = a.transpose(2, 0, 1).astype(np.float32) / np.float32(255)  # ToTensor: .div(255)
    return (t - np.float32(0.5)) / np.float32(0.5)                 # Normalize: .sub_(mean).div_(std)


def cloth_mask(cloth_chw, threshold=240):
    """get_input_cloth_mask (tryon_dataset.py:168-175).  Note the reference compares the NORMALISED cloth (values in
    [-1, 1]) with the 0-255 threshold, so with the default 240 the mask is all ones; reproduced as written."""
    m = np.where(cloth_chw >= np.float32(threshold), np.float32(0), np.float32(1))
    return m[0:1]


def person_head(im_chw, parse_hw):
    """get_person_head (tryon_dataset.py:323-344): im * phead - (1 - phead)."""
    phead = np.zeros(parse_hw.shape, np.float32)
    for lab in LIP_HEAD_LABELS:
        phead += (parse_hw == lab).astype(np.float32)
    return im_chw * phead - (np.float32(1) - phead)


def pil_resize_coeffs(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle, support 1) filter.
    Returns (bounds [out,2] int32 = (xmin, count), kk [out, ksize] int32 fixed-point weights)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size  # (double)(in1 - in0) / outSize, in0/in1 are floats
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(xmax, np.float64)
        ww = 0.0
        for x in range(xmax):
            t = (x + xmin - center + 0.5) * ss
            if t < 0.0:
                t = -t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)  # C cast: truncation toward zero
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis_u8(img, bounds, kk, axis):
    """One 8-bit pass of ImagingResampleHorizontal_8bpc / Vertical_8bpc: ss = 1 << (PB - 1); ss += px * k; clip8(ss >> PB)."""
    a = img.astype(np.int64)
    out_size = bounds.shape[0]
    shape = list(a.shape)
    shape[axis] = out_size
    out = np.zeros(shape, np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        k = kk[xx, :cnt].astype(np.int64)
        seg = a[:, xmin:xmin + cnt] if axis == 1 else a[xmin:xmin + cnt, :]
        acc = (1 << (PRECISION_BITS - 1)) + (seg * (k[None, :] if axis == 1 else k[:, None])).sum(axis=axis)
        v = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, xx] = v
        else:
            out[xx, :] = v
    return out


def pil_resize_bilinear_u8(img_hw, out_w, out_h):
    """Image.resize((out_w, out_h), Image.BILINEAR) on an 8-bit single-band image: horizontal pass, then vertical
    (ImagingResample, Resample.c), each rounding to uint8."""
    h, w = img_hw.shape
    tmp = img_hw
    if out_w != w:
        b, k = pil_resize_coeffs(w, out_w)
        tmp = _resample_axis_u8(tmp, b, k, axis=1)
    if out_h != h:
        b, k = pil_resize_coeffs(h, out_h)
        tmp = _resample_axis_u8(tmp, b, k, axis=0)
    return tmp


def body_silhouette_u8(parse_hw):
    """The 8-bit image inside get_person_body_silhouette (tryon_dataset.py:346-358): (parse > 0) * 255, downsampled
    16x and upsampled back, both PIL BILINEAR."""
    h, w = parse_hw.shape
    shape = ((parse_hw > 0).astype(np.float32) * 255).astype(np.uint8)
    small = pil_resize_bilinear_u8(shape, w // 16, h // 16)
    return pil_resize_bilinear_u8(small, w, h)


def body_silhouette(parse_hw):
    """get_person_body_silhouette (tryon_dataset.py:346-367): normalised [1,H,W] f32."""
    return norm_u8(body_silhouette_u8(parse_hw))


def cocopose_vis_u8(pose_data, h, w, radius=5):
    """The visualisation image of convert_pose_data_to_pose_map_and_vis (tryon_dataset.py:406-437): union of the filled
    (2r+1)-squares PIL's ImageDraw.rectangle draws around every joint with x > 1 and y > 1 (inclusive corners,
    coordinates truncated to int by the C drawing code, clipped to the image)."""
    im = np.zeros((h, w), np.uint8)
    if pose_data is None:
        return im
    for x, y, _ in np.asarray(pose_data, dtype=np.float64).reshape(-1, 3):
        if x > 1 and y > 1:
            x0, y0, x1, y1 = int(x - radius), int(y - radius), int(x + radius), int(y + radius)
            xa, xb = max(x0, 0), min(x1, w - 1)
            ya, yb = max(y0, 0), min(y1, h - 1)
            if xa <= xb and ya <= yb:
                im[ya:yb + 1, xa:xb + 1] = 255
    return im


def cocopose(pose_data, h, w, radius=5, n_joints=18):
    """convert_pose_data_to_pose_map_and_vis (tryon_dataset.py:389-447) -> (pose_map [J,H,W], im_cocopose [1,H,W]).
    As written in the reference, pose_map[i] is filled from the blank image BEFORE the square is drawn
    (tryon_dataset.py:415-423), so every channel of the map the networks consume is the constant -1; only the
    visualisation carries the squares.  Reproduced as written."""
    j = n_joints if pose_data is None else np.asarray(pose_data).reshape(-1, 3).shape[0]
    pose_map = np.full((j, h, w), -1.0, np.float32)
    return pose_map, norm_u8(cocopose_vis_u8(pose_data, h, w, radius))


def decode_flo(buf):
    """flownet2_pytorch/utils/flow_utils.py:7-26 readFlow + permute + Normalize(0.5, 0.5) (tryon_dataset.py:121,288-289):
    Middlebury .flo bytes (magic 202021.25, int32 w, int32 h, interleaved f32 u,v) -> f32 [2,h,w] = (x - 0.5) / 0.5."""
    b = np.frombuffer(buf, dtype=np.uint8)
    magic = b[:4].view(np.float32)[0]
    if magic != np.float32(202021.25):
        raise ValueError("Magic number incorrect. Invalid .flo file")
    w, h = int(b[4:8].view(np.int32)[0]), int(b[8:12].view(np.int32)[0])
    data = b[12:12 + 8 * w * h].view(np.float32).reshape(h, w, 2)
    t = data.transpose(2, 0, 1)
    return (t - np.float32(0.5)) / np.float32(0.5)


def frame_prep(image_hwc, parse_hw, cloth_hwc, densepose_hwc, pose_data, cloth_mask_threshold=240, radius=5):
    """One frame of TryonDataset.get_person_representation + get_cloth_representation
    (tryon_dataset.py:156-166,203-251) from decoded 8-bit images.  Returns the batch keys the try-on stages read."""
    h, w = parse_hw.shape
    image = norm_u8(image_hwc)
    cloth = norm_u8(cloth_hwc)
    pose_map, im_cocopose = cocopose(pose_data, h, w, radius)
    return {
        "image": image, "cloth": cloth, "cloth_mask": cloth_mask(cloth, cloth_mask_threshold),
        "densepose": norm_u8(densepose_hwc), "silhouette": body_silhouette(parse_hw),
        "im_head": person_head(image, parse_hw),
        "agnostic": np.concatenate([body_silhouette(parse_hw), person_head(image, parse_hw)], 0),
        "cocopose": pose_map, "im_cocopose": im_cocopose,
    }


def synth_frame(seed, h=256, w=192):
    """Seeded synthetic 8-bit inputs with the structure of a VVT sample: photo-like images, a blocky LIP parse map with
    a background margin, 18 key points (some missing = zeros, some fractional / near the border)."""
    r = np.random.RandomState(seed)
    image = r.randint(0, 256, (h, w, 3), dtype=np.uint8)
    cloth = r.randint(0, 256, (h, w, 3), dtype=np.uint8)
    cloth[: h // 8] = 255  # white background rows (values above the nominal threshold)
    densepose = r.randint(0, 256, (h, w, 3), dtype=np.uint8)
    labels = r.randint(0, 20, (h // 16 + 1, w // 8 + 1)).astype(np.uint8)
    parse = np.kron(labels, np.ones((16, 8), np.uint8))[3:3 + h, 5:5 + w].copy()
    parse[:, : w // 6] = 0
    parse[: h // 10] = 0
    pose = np.zeros((18, 3), np.float64)
    pose[:, 0] = r.uniform(-4, w + 4, 18)
    pose[:, 1] = r.uniform(-4, h + 4, 18)
    pose[:, 2] = r.uniform(0, 1, 18)
    pose[r.randint(0, 18, 4)] = 0.0
    pose[0, :2] = (1.0, 50.0)       # x == 1: not drawn (tryon_dataset.py:426)
    pose[1, :2] = (w - 1.5, 2.25)   # square clipped at two borders
    return image, parse, cloth, densepose, pose
