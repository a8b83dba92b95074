"""CPU oracle for the GMM (warp) stage.  Reference: models/networks/cpvton/warp.py, models/warp_model.py."""
import numpy as np
import torch
import torch.nn.functional as F


def _bn(x, sd, p):
    """nn.BatchNorm2d in eval mode (running statistics)."""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, 1e-5)


def feature_extraction(sd, prefix, x, n_layers=3):
    """FeatureExtraction.forward (warp.py:9-36): [conv4x4s2, ReLU, BN] x (1+n_layers), conv3x3, ReLU, BN, conv3x3, ReLU.
    Sequential indices: 0 conv,1 relu,2 bn | 3,4,5 | 6,7,8 | 9,10,11 | 12 conv3,13 relu,14 bn | 15 conv3,16 relu."""
    p = prefix + "model."
    i = 0
    for _ in range(1 + n_layers):
        x = F.conv2d(x, sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"], stride=2, padding=1)
        x = F.relu(x)
        x = _bn(x, sd, f"{p}{i + 2}.")
        i += 3
    x = F.conv2d(x, sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"], stride=1, padding=1)
    x = F.relu(x)
    x = _bn(x, sd, f"{p}{i + 2}.")
    i += 3
    x = F.conv2d(x, sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"], stride=1, padding=1)
    return F.relu(x)


def feature_l2norm(f):
    """FeatureL2Norm.forward (warp.py:43-50)."""
    norm = torch.pow(torch.sum(torch.pow(f, 2), 1) + 1e-6, 0.5).unsqueeze(1).expand_as(f)
    return torch.div(f, norm)


def feature_correlation(fa, fb):
    """FeatureCorrelation.forward (warp.py:57-67): out[b, wA*h+hA, hB, wB] = <A[:,hA,wA], B[:,hB,wB]>."""
    b, c, h, w = fa.size()
    fa = fa.transpose(2, 3).contiguous().view(b, c, h * w)
    fb = fb.view(b, c, h * w).transpose(1, 2)
    mul = torch.bmm(fb, fa)
    return mul.view(b, h, w, h * w).transpose(2, 3).transpose(1, 2)


def feature_regression(sd, prefix, x):
    """FeatureRegression.forward (warp.py:70-99): conv(0) bn(1) relu | conv(3) bn(4) | conv(6) bn(7) | conv(9) bn(10)
    -> flatten (NCHW) -> linear -> tanh."""
    p = prefix + "conv."
    for i, (s, pad) in zip((0, 3, 6, 9), ((2, 1), (2, 1), (1, 1), (1, 1))):
        x = F.conv2d(x, sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"], stride=s, padding=pad)
        x = F.relu(_bn(x, sd, f"{p}{i + 1}."))
    x = x.contiguous().view(x.size(0), -1)
    x = F.linear(x, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    return torch.tanh(x)


class TpsTables:
    """Constants of TpsGridGen.__init__ / compute_L_inverse (warp.py:116-189), same numpy/torch calls."""

    def __init__(self, out_h=256, out_w=192, grid_size=3):
        self.out_h, self.out_w, self.grid_size = out_h, out_w, grid_size
        gx, gy = np.meshgrid(np.linspace(-1, 1, out_w), np.linspace(-1, 1, out_h))
        self.grid_X = torch.FloatTensor(gx)  # [H,W]
        self.grid_Y = torch.FloatTensor(gy)
        axis = np.linspace(-1, 1, grid_size)
        self.N = grid_size * grid_size
        P_Y, P_X = np.meshgrid(axis, axis)
        self.P_X = torch.FloatTensor(np.reshape(P_X, (-1, 1)))  # [N,1]
        self.P_Y = torch.FloatTensor(np.reshape(P_Y, (-1, 1)))
        self.Li = self._l_inverse(self.P_X, self.P_Y)  # [N+3,N+3]

    @staticmethod
    def _l_inverse(X, Y):
        N = X.size(0)
        Xm, Ym = X.expand(N, N), Y.expand(N, N)
        d2 = torch.pow(Xm - Xm.transpose(0, 1), 2) + torch.pow(Ym - Ym.transpose(0, 1), 2)
        d2[d2 == 0] = 1
        K = torch.mul(d2, torch.log(d2))
        O = torch.FloatTensor(N, 1).fill_(1)
        Z = torch.FloatTensor(3, 3).fill_(0)
        P = torch.cat((O, X, Y), 1)
        L = torch.cat((torch.cat((K, P), 1), torch.cat((P.transpose(0, 1), Z), 1)), 0)
        return torch.inverse(L)


def tps_grid(theta, t: TpsTables):
    """TpsGridGen.apply_transformation (warp.py:191-318) without the [B,H,W,1,N] temporaries; same arithmetic
    (W = Li[:N,:N] Q, A = Li[N:,:N] Q, U = d^2 log d^2 with d^2==0 -> 1)."""
    B, N = theta.shape[0], t.N
    Q_X = theta[:, :N].unsqueeze(2) + t.P_X.unsqueeze(0)  # [B,N,1]
    Q_Y = theta[:, N:].unsqueeze(2) + t.P_Y.unsqueeze(0)
    Li = t.Li.unsqueeze(0)
    W_X = torch.bmm(Li[:, :N, :N].expand(B, N, N), Q_X)  # [B,N,1]
    W_Y = torch.bmm(Li[:, :N, :N].expand(B, N, N), Q_Y)
    A_X = torch.bmm(Li[:, N:, :N].expand(B, 3, N), Q_X)  # [B,3,1]
    A_Y = torch.bmm(Li[:, N:, :N].expand(B, 3, N), Q_Y)
    gx, gy = t.grid_X, t.grid_Y  # [H,W]
    dx = gx.unsqueeze(2) - t.P_X.view(1, 1, N)
    dy = gy.unsqueeze(2) - t.P_Y.view(1, 1, N)
    d2 = torch.pow(dx, 2) + torch.pow(dy, 2)
    d2[d2 == 0] = 1
    U = torch.mul(d2, torch.log(d2))  # [H,W,N]
    xs = (A_X[:, 0].view(B, 1, 1) + A_X[:, 1].view(B, 1, 1) * gx + A_X[:, 2].view(B, 1, 1) * gy
          + torch.sum(W_X.view(B, 1, 1, N) * U.unsqueeze(0), 3))
    ys = (A_Y[:, 0].view(B, 1, 1) + A_Y[:, 1].view(B, 1, 1) * gx + A_Y[:, 2].view(B, 1, 1) * gy
          + torch.sum(W_Y.view(B, 1, 1, N) * U.unsqueeze(0), 3))
    return torch.stack((xs, ys), 3)  # [B,H,W,2]


def gmm_forward(sd, inputA, inputB, t: TpsTables):
    """WarpModel.forward (warp_model.py:63-72) -> (grid, theta)."""
    fa = feature_l2norm(feature_extraction(sd, "extractionA.", inputA))
    fb = feature_l2norm(feature_extraction(sd, "extractionB.", inputB))
    corr = feature_correlation(fa, fb)
    theta = feature_regression(sd, "regression.", corr)
    return tps_grid(theta, t), theta


def grid_sample(x, grid, padding_mode):
    """The F.grid_sample call sites (warp_model.py:85-86,143-145): bilinear, align_corners=False."""
    return F.grid_sample(x, grid, mode="bilinear", padding_mode=padding_mode, align_corners=False)
