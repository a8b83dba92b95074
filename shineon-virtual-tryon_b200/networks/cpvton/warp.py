"""GMM networks (reference: models/networks/cpvton/warp.py:9-318): FeatureExtraction, FeatureL2Norm,
FeatureCorrelation, FeatureRegression, TpsGridGen — same ctor signatures, module trees and state_dict keys.

Engine (eval mode): every conv is the tcgen05 implicit-GEMM kernel with bias / ReLU / folded BatchNorm in
its epilogue, writing bf16 hi/lo NHWC planes straight into the next conv; L2-norm + all-pairs correlation is
one fp32 tiled kernel; Linear+tanh one small kernel; the TPS grid is generated (and optionally consumed by
grid_sample) in one gather kernel.
"""
import numpy as np
import torch
from torch import nn

from .. import init_weights
from ... import ops
from .._engine_util import fold_bn, params_signature, require_cuda


def _nhwc_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


class FeatureExtraction(nn.Module):
    def __init__(self, input_nc, ngf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        downconv = nn.Conv2d(input_nc, ngf, kernel_size=4, stride=2, padding=1)
        model = [downconv, nn.ReLU(True), norm_layer(ngf)]
        for i in range(n_layers):
            in_ngf = 2 ** i * ngf if 2 ** i * ngf < 512 else 512
            out_ngf = 2 ** (i + 1) * ngf if 2 ** i * ngf < 512 else 512
            downconv = nn.Conv2d(in_ngf, out_ngf, kernel_size=4, stride=2, padding=1)
            model += [downconv, nn.ReLU(True)]
            model += [norm_layer(out_ngf)]
        model += [nn.Conv2d(512, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(True)]
        model += [norm_layer(512)]
        model += [nn.Conv2d(512, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(True)]
        self.model = nn.Sequential(*model)
        init_weights(self.model, init_type="normal")
        self._packed = None
        self.precision = None

    def _layers(self, prec):
        sig = (params_signature(self), prec)
        if self._packed is not None and self._packed[0] == sig:
            return self._packed[1]
        require_cuda(self, "FeatureExtraction")
        mods = list(self.model)
        layers = []
        i = 0
        while i < len(mods):
            conv = mods[i]
            assert isinstance(conv, nn.Conv2d) and isinstance(mods[i + 1], nn.ReLU)
            bn = mods[i + 2] if i + 2 < len(mods) else None
            if bn is not None and not isinstance(bn, nn.BatchNorm2d):
                raise NotImplementedError(f"FeatureExtraction norm {type(bn).__name__} has no native kernel")
            sc, sh = fold_bn(bn) if bn is not None else (None, None)
            if i == 0 and conv.in_channels <= 32:
                # tiny Cin: im2col'd input + dense 1x1 GEMM instead of one mostly-zero K-block per tap
                i2c = ops.first_layer_conv(conv.weight, conv.bias, conv.stride[0], conv.padding[0], prec=prec)
                layers.append((i2c.pc, sc, sh, i2c))
            else:
                pc = ops.PackedConv(conv.weight, conv.bias, stride=conv.stride[0], pad=conv.padding[0], prec=prec)
                layers.append((pc, sc, sh, None))
            i += 3
        self._packed = (sig, layers)
        return layers

    def forward_nhwc(self, x):
        """x f32 NCHW -> f32 NHWC features (conv -> ReLU -> BN ... conv -> ReLU, warp.py:13-31).
        x may also be the stem's operand built elsewhere (ops.frame_prep_planes): an ops.S2dInput for the space-to-depth
        stem, or the im2col Planes for the 3-channel stem."""
        prec = ops.resolve_precision(self.precision)
        layers = self._layers(prec)
        first = layers[0][3]
        prebuilt = None
        if isinstance(x, ops.S2dInput):
            assert isinstance(first, ops.S2dConv) and x.C == first.Cin, "stem operand does not match this network's first layer"
            prebuilt = x.planes
        elif isinstance(x, ops.Planes):
            assert isinstance(first, ops.Im2colConv) and x.cpad == first.pc.cin_pad, "stem operand does not match the first layer"
            prebuilt = x
        else:
            x = x.contiguous()
        a = None if first is not None else ops.nchw_to_planes(x, prec=prec)
        for li, (pc, sc, sh, i2c) in enumerate(layers):
            last = li == len(layers) - 1
            if i2c is not None and prebuilt is not None:
                f32, a = ops.conv2d(prebuilt, i2c.pc, scale=sc, shift=sh, pre_act="relu", want_f32=last, want_planes=not last)
            elif i2c is not None:  # tiny-Cin stem: layout pass (space-to-depth / im2col planes) + GEMM
                f32, a = i2c.conv(x, scale=sc, shift=sh, pre_act="relu", want_f32=last, want_planes=not last)
            else:
                f32, a = ops.conv2d(a, pc, scale=sc, shift=sh, pre_act="relu", want_f32=last, want_planes=not last)
        return f32

    def forward(self, x):
        return _nhwc_to_nchw(self.forward_nhwc(x))


class FeatureL2Norm(nn.Module):
    """x / sqrt(sum_c x^2 + 1e-6) (warp.py:39-50).  WarpModel fuses it into the correlation kernel
    (FeatureCorrelation.forward_fused); standalone it is one small kernel on the reference layout."""

    def forward(self, feature):
        return ops.feature_l2norm(feature.contiguous())


class FeatureCorrelation(nn.Module):
    def forward_fused(self, fa_nhwc, fb_nhwc, prec=None, want_f32=False):
        """FeatureL2Norm x2 + FeatureCorrelation (warp.py:39-67) on f32 NHWC features: a per-image tensor-core GEMM over
        the normalised planes where the channel count allows (ops.l2norm_correlation_tc), else the fp32 kernel."""
        if ops.CORRELATION_TC and fa_nhwc.shape[-1] % 64 == 0:
            return ops.l2norm_correlation_tc(fa_nhwc, fb_nhwc, prec=prec, want_f32=want_f32)
        return ops.l2norm_correlation(fa_nhwc, fb_nhwc, want_f32=want_f32, want_planes=True, prec=prec)

    def forward(self, feature_A, feature_B):
        """warp.py:53-67 on (already normalised) NCHW features: [B,C,h,w] x2 -> [B,h*w,h,w],
        out[b, wA*h + hA, hB, wB] = sum_c A[b,c,hA,wA] * B[b,c,hB,wB]."""
        fa = feature_A.permute(0, 2, 3, 1).contiguous()
        fb = feature_B.permute(0, 2, 3, 1).contiguous()
        corr, _ = ops.l2norm_correlation(fa, fb, want_f32=True, want_planes=False, normalize=False)
        return corr.permute(0, 3, 1, 2).contiguous()


class FeatureRegression(nn.Module):
    def __init__(self, input_nc=512, output_dim=6):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(input_nc, 512, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(512), nn.ReLU(inplace=True),
            nn.Conv2d(512, 256, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            nn.Conv2d(256, 128, kernel_size=3, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
            nn.Conv2d(128, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
        )
        self.linear = nn.Linear(64 * 4 * 3, output_dim)
        self.tanh = nn.Tanh()
        self._packed = None
        self.precision = None

    def _layers(self, prec):
        sig = (params_signature(self), prec)
        if self._packed is not None and self._packed[0] == sig:
            return self._packed[1]
        require_cuda(self, "FeatureRegression")
        mods = list(self.conv)
        layers = []
        for i in range(0, len(mods), 3):
            conv, bn = mods[i], mods[i + 1]
            sc, sh = fold_bn(bn)
            layers.append((ops.PackedConv(conv.weight, conv.bias, stride=conv.stride[0], pad=conv.padding[0],
                                          prec=prec), sc, sh))
        self._packed = (sig, layers)
        return layers

    def forward_planes(self, corr_planes):
        """corr_planes: Planes [B,16,12,192] -> theta [B, output_dim] (warp.py:94-99)."""
        layers = self._layers(corr_planes.prec)
        a = corr_planes
        for li, (pc, sc, sh) in enumerate(layers):
            last = li == len(layers) - 1
            f32, a = ops.conv2d(a, pc, scale=sc, shift=sh, post_act="relu", want_f32=last, want_planes=not last)
        return ops.linear_tanh(f32, self.linear.weight.detach(), self.linear.bias.detach())

    def forward(self, x):
        return self.forward_planes(ops.nchw_to_planes(x.contiguous(), prec=ops.resolve_precision(self.precision)))


class TpsGridGen(nn.Module):
    def __init__(self, out_h=256, out_w=192, use_regular_grid=True, grid_size=3, reg_factor=0):
        super().__init__()
        self.out_h, self.out_w = out_h, out_w
        self.reg_factor = reg_factor
        self.grid_size = grid_size
        # same numpy / torch calls as the reference constructor (warp.py:124-157) so the constants are identical
        self.grid = np.zeros([self.out_h, self.out_w, 3], dtype=np.float32)
        self.grid_X, self.grid_Y = np.meshgrid(np.linspace(-1, 1, out_w), np.linspace(-1, 1, out_h))
        self.grid_X = torch.FloatTensor(self.grid_X).unsqueeze(0).unsqueeze(3)
        self.grid_Y = torch.FloatTensor(self.grid_Y).unsqueeze(0).unsqueeze(3)
        if use_regular_grid:
            axis_coords = np.linspace(-1, 1, grid_size)
            self.N = grid_size * grid_size
            P_Y, P_X = np.meshgrid(axis_coords, axis_coords)
            P_X = torch.FloatTensor(np.reshape(P_X, (-1, 1)))
            P_Y = torch.FloatTensor(np.reshape(P_Y, (-1, 1)))
            self.P_X_base = P_X.clone()
            self.P_Y_base = P_Y.clone()
            self.Li = self.compute_L_inverse(P_X, P_Y).unsqueeze(0)
            self.P_X = P_X.unsqueeze(2).unsqueeze(3).unsqueeze(4).transpose(0, 4)
            self.P_Y = P_Y.unsqueeze(2).unsqueeze(3).unsqueeze(4).transpose(0, 4)
        else:
            raise NotImplementedError("only the regular control grid is used by the reference")
        self._tables = {}

    def compute_L_inverse(self, X, Y):
        N = X.size()[0]
        Xmat, Ymat = X.expand(N, N), Y.expand(N, N)
        P_dist_squared = torch.pow(Xmat - Xmat.transpose(0, 1), 2) + torch.pow(Ymat - Ymat.transpose(0, 1), 2)
        P_dist_squared[P_dist_squared == 0] = 1
        K = torch.mul(P_dist_squared, torch.log(P_dist_squared))
        O = torch.FloatTensor(N, 1).fill_(1)
        Z = torch.FloatTensor(3, 3).fill_(0)
        P = torch.cat((O, X, Y), 1)
        L = torch.cat((torch.cat((K, P), 1), torch.cat((P.transpose(0, 1), Z), 1)), 0)
        return torch.inverse(L)

    def tables(self, device):
        key = str(device)
        if key not in self._tables:
            self._tables[key] = ops.TpsTablesDev(self.Li[0], self.P_X_base, self.P_Y_base, self.grid_X[0, 0, :, 0],
                                                 self.grid_Y[0, :, 0, 0], self.grid_size, device)
        return self._tables[key]

    def forward(self, theta):
        """theta [B, 2*grid_size^2] -> sampling grid [B,H,W,2] (warp.py:159-167)."""
        if theta.dim() == 4:
            theta = theta.reshape(theta.shape[0], -1)
        return ops.tps_grid(theta.contiguous(), self.tables(theta.device), self.out_h, self.out_w)

    def warp_u8(self, theta, cloth_u8, unet_in):
        """Fused TPS + grid_sample(border) of the decoded 8-bit cloth [B,H,W,3]: returns the f32 NCHW warped cloth and
        writes the same samples into the cloth' channels of the U-Net stem operand (ops.tps_warp_u8_planes)."""
        return ops.tps_warp_u8_planes(theta.contiguous(), self.tables(theta.device), cloth_u8, unet_in)

    def warp(self, theta, inputs, want_grid=False):
        """Fused TPS + grid_sample: inputs = [(tensor [B,C,H,W], padding_mode), ...] (<= 3)."""
        return ops.tps_grid_sample(theta.contiguous(), self.tables(theta.device), self.out_h, self.out_w, inputs,
                                   want_grid=want_grid)
