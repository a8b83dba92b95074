"""CPU oracle for the three CUDA-only reference ops (vectorised torch, float32 unless noted).

Reference kernels (models/flownet2_pytorch/networks/...):
  resample2d_package/resample2d_kernel.cu   :16-72 forward, :76-125 backward(input1), :128-198 backward(flow)
  channelnorm_package/channelnorm_kernel.cu :19-60 forward, :64-96 backward
  correlation_package/correlation_cuda_kernel.cu :74-147 forward, :151-334 backward;  shapes correlation_cuda.cc:19-38

PARITY NOTE: these kernels have no CPU implementation in the reference and the build container has no GPU,
so this restatement is pinned against the reference only via oracle/_ref (the reference's own .cu files
compiled for sm_100, run on the GPU box by tests/test_ref_ext_gpu.py) when that build succeeds.
"""
import math

import torch


# ------------------------------------------------------------------------------------ Resample2d
def _resample_indices(flow, H, W):
    B = flow.shape[0]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    xf = xs.unsqueeze(0) + flow[:, 0]  # resample2d_kernel.cu:37-41
    yf = ys.unsqueeze(0) + flow[:, 1]
    fx, fy = torch.floor(xf), torch.floor(yf)
    xL = fx.clamp(0, W - 1).long()  # :45-48  max(min(int(floor(xf)), W-1), 0)
    xR = (fx + 1).clamp(0, W - 1).long()
    yT = fy.clamp(0, H - 1).long()
    yB = (fy + 1).clamp(0, H - 1).long()
    return xf, yf, fx, fy, xL, xR, yT, yB


def _gather(img, y, x):
    B, C, H, W = img.shape
    idx = (y * W + x).view(B, 1, -1).expand(B, C, -1)
    return img.reshape(B, C, H * W).gather(2, idx).view(B, C, *y.shape[1:])


def resample2d_fwd(in1, flow, kernel_size=1, bilinear=True):
    """kernel_resample2d_update_output (resample2d_kernel.cu:16-72), kernel_size == 1."""
    assert kernel_size == 1
    B, C, H, W = in1.shape[0], in1.shape[1], flow.shape[2], flow.shape[3]
    xf, yf, fx, fy, xL, xR, yT, yB = _resample_indices(flow, H, W)
    if not bilinear:
        xN = torch.floor(xf + 0.5).clamp(0, W - 1).long()  # :66-69
        yN = torch.floor(yf + 0.5).clamp(0, H - 1).long()
        return _gather(in1, yN, xN)
    alpha = (xf - fx).unsqueeze(1)  # :42-43
    beta = (yf - fy).unsqueeze(1)
    val = torch.zeros(B, C, H, W)
    val = val + (1.0 - alpha) * (1.0 - beta) * _gather(in1, yT, xL)  # :56-59, same order
    val = val + alpha * (1.0 - beta) * _gather(in1, yT, xR)
    val = val + (1.0 - alpha) * beta * _gather(in1, yB, xL)
    val = val + alpha * beta * _gather(in1, yB, xR)
    return val


def resample2d_bwd(in1, flow, grad_out):
    """kernel_resample2d_backward_input1 (:76-125) + kernel_resample2d_backward_input2 (:128-198).
    NB alpha/beta of the input1 gradient use truncation `xf - int(xf)` (:105-106), not floor."""
    B, C, H, W = in1.shape
    xf, yf, fx, fy, xL, xR, yT, yB = _resample_indices(flow, H, W)
    alpha = (xf - torch.trunc(xf)).unsqueeze(1)
    beta = (yf - torch.trunc(yf)).unsqueeze(1)
    g1 = torch.zeros(B, C, H * W)

    def scatter(y, x, wgt):
        idx = (y * W + x).view(B, 1, -1).expand(B, C, -1)
        g1.scatter_add_(2, idx, (wgt * grad_out).reshape(B, C, -1))

    scatter(yT, xL, (1 - alpha) * (1 - beta))  # :118-121
    scatter(yT, xR, alpha * (1 - beta))
    scatter(yB, xL, (1 - alpha) * beta)
    scatter(yB, xR, alpha * beta)
    g1 = g1.view(B, C, H, W)
    vTL, vTR = _gather(in1, yT, xL), _gather(in1, yT, xR)
    vBL, vBR = _gather(in1, yB, xL), _gather(in1, yB, xR)
    gx_gamma = (1 - (yf - fy)).unsqueeze(1)  # even channel (:181-192)
    gy_gamma = (1 - (xf - fx)).unsqueeze(1)  # odd channel (:168-179)
    gdx = (gx_gamma * grad_out * (vTR - vTL) + (1 - gx_gamma) * grad_out * (vBR - vBL)).sum(1)
    gdy = (gy_gamma * grad_out * (vBL - vTL) + (1 - gy_gamma) * grad_out * (vBR - vTR)).sum(1)
    return g1, torch.stack((gdx, gdy), 1)


# ------------------------------------------------------------------------------------ ChannelNorm
def channelnorm_fwd(x):
    """kernel_channelnorm_update_output (channelnorm_kernel.cu:19-60): sqrt(sum_c x^2); norm_deg unused."""
    return torch.sqrt((x * x).sum(1, keepdim=True))


def channelnorm_bwd(x, out, grad_out):
    """kernel_channelnorm_backward_input1 (:64-96): gOut * in / (out + 1e-9), the division in double."""
    return ((grad_out * x).double() / (out.double() + 1e-9)).float()


# ------------------------------------------------------------------------------------ Correlation
def correlation_out_shape(C, H, W, pad, k, maxd, s1, s2):
    """correlation_cuda.cc:19-38."""
    kr = (k - 1) // 2
    border = kr + maxd
    D = (maxd // s2) * 2 + 1
    oh = int(math.ceil(float(H + 2 * pad - 2 * border) / float(s1)))
    ow = int(math.ceil(float(W + 2 * pad - 2 * border) / float(s1)))
    return D * D, oh, ow


def correlation_fwd(in1, in2, pad, k, maxd, s1, s2):
    """correlation_forward (correlation_cuda_kernel.cu:74-147) on zero-padded inputs (channels_first, :47-70)."""
    B, C, H, W = in1.shape
    kr = (k - 1) // 2
    drad = maxd // s2
    D = 2 * drad + 1
    oc, oh, ow = correlation_out_shape(C, H, W, pad, k, maxd, s1, s2)
    # extra margin so that every window the reference would read from its padded buffer exists here too
    m = maxd + kr + s2 * drad
    p1 = torch.nn.functional.pad(in1, (pad + m, pad + m, pad + m, pad + m))
    p2 = torch.nn.functional.pad(in2, (pad + m, pad + m, pad + m, pad + m))
    out = torch.zeros(B, oc, oh, ow)
    ys = torch.arange(oh) * s1 + maxd + m  # y1 in padded coords (+m margin)
    xs = torch.arange(ow) * s1 + maxd + m
    nelems = k * k * C
    for tj in range(-drad, drad + 1):
        for ti in range(-drad, drad + 1):
            acc = torch.zeros(B, oh, ow)
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    a = p1[:, :, (ys + j)[:, None], (xs + i)[None, :]]
                    b = p2[:, :, (ys + tj * s2 + j)[:, None], (xs + ti * s2 + i)[None, :]]
                    acc = acc + (a * b).sum(1)
            out[:, (tj + drad) * D + (ti + drad)] = acc / nelems
    return out


def correlation_bwd(in1, in2, grad_out, pad, k, maxd, s1, s2):
    """correlation_backward_input1/2 (correlation_cuda_kernel.cu:151-334), loops as in the kernels
    (C++ integer division truncating toward zero)."""
    B, C, H, W = in1.shape
    kr = (k - 1) // 2
    drad = maxd // s2
    D = 2 * drad + 1
    oc, oh, ow = correlation_out_shape(C, H, W, pad, k, maxd, s1, s2)
    nelems = float(k * k * C)
    m = maxd + kr + s2 * drad
    p1 = torch.nn.functional.pad(in1, (m, m, m, m))
    p2 = torch.nn.functional.pad(in2, (m, m, m, m))
    tdiv = lambda a, b: int(a / b)  # trunc toward zero
    g1 = torch.zeros_like(in1)
    g2 = torch.zeros_like(in2)
    for yi in range(H):
        for xi in range(W):
            y, x = yi + pad, xi + pad
            # ---- input1 (:170-240)
            xmin, ymin = tdiv(x - kr - maxd, s1), tdiv(y - kr - maxd, s1)
            xmax, ymax = tdiv(x + kr - maxd, s1), tdiv(y + kr - maxd, s1)
            if not (xmax < 0 or ymax < 0 or xmin >= ow or ymin >= oh or xmin > xmax or ymin > ymax):
                xmin, xmax = max(0, xmin), min(ow - 1, xmax)
                ymin, ymax = max(0, ymin), min(oh - 1, ymax)
                for tc in range(oc):
                    i2, j2 = (tc % D - drad) * s2, (tc // D - drad) * s2
                    v2 = p2[:, :, yi + j2 + m, xi + i2 + m]  # [B,C]
                    gs = grad_out[:, tc, ymin:ymax + 1, xmin:xmax + 1].sum((1, 2))  # [B]
                    g1[:, :, yi, xi] += gs[:, None] * v2
            # ---- input2 (:263-333)
            for tc in range(oc):
                i2, j2 = (tc % D - drad) * s2, (tc // D - drad) * s2
                xmin, ymin = tdiv(x - kr - maxd - i2, s1), tdiv(y - kr - maxd - j2, s1)
                xmax, ymax = tdiv(x + kr - maxd - i2, s1), tdiv(y + kr - maxd - j2, s1)
                if xmax < 0 or ymax < 0 or xmin >= ow or ymin >= oh or xmin > xmax or ymin > ymax:
                    continue
                xmin, xmax = max(0, xmin), min(ow - 1, xmax)
                ymin, ymax = max(0, ymin), min(oh - 1, ymax)
                v1 = p1[:, :, yi - j2 + m, xi - i2 + m]
                gs = grad_out[:, tc, ymin:ymax + 1, xmin:xmax + 1].sum((1, 2))
                g2[:, :, yi, xi] += gs[:, None] * v1
    return g1 / nelems, g2 / nelems
