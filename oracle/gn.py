"""CPU restatement of the tower normalisation -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``nn.GroupNorm(G, C)`` followed by ``nn.ReLU`` as the reference's towers apply them
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:644-675, :733-736): torch semantics --
statistics over (C/G, H, W) per image, biased variance, eps inside the square root, per-channel affine, then
max(., 0).  float64 numpy; pinned against ``torch.nn.functional.group_norm`` + ``relu`` (forward and autograd
gradients) in tests/test_oracle_gn.py -- torch IS the reference's implementation of these two layers.
"""
import numpy as np


def forward(x, gamma, beta, G, eps=1e-5, relu=True):
    x = np.asarray(x, np.float64)
    N, C, H, W = x.shape
    xg = x.reshape(N, G, -1)
    mean = xg.mean(axis=2, keepdims=True)
    var = xg.var(axis=2, keepdims=True)          # biased
    rstd = 1.0 / np.sqrt(var + eps)
    xh = ((xg - mean) * rstd).reshape(N, C, H, W)
    y = xh * np.asarray(gamma, np.float64)[None, :, None, None] + np.asarray(beta, np.float64)[None, :, None, None]
    if relu:
        y = np.maximum(y, 0.0)
    return y, mean.reshape(N, G), rstd.reshape(N, G)


def backward(x, gamma, beta, G, grad_y, eps=1e-5, relu=True):
    """-> dict(grad_x, grad_gamma, grad_beta)"""
    x = np.asarray(x, np.float64)
    gy = np.asarray(grad_y, np.float64)
    gamma = np.asarray(gamma, np.float64)
    N, C, H, W = x.shape
    y, mean, rstd = forward(x, gamma, beta, G, eps, relu)
    if relu:
        gy = gy * (y > 0)
    xh = ((x.reshape(N, G, -1) - mean[:, :, None]) * rstd[:, :, None]).reshape(N, C, H, W)
    grad_gamma = (gy * xh).sum(axis=(0, 2, 3))
    grad_beta = gy.sum(axis=(0, 2, 3))
    gyg = gy * gamma[None, :, None, None]
    m = (C // G) * H * W
    s1 = gyg.reshape(N, G, -1).sum(axis=2) / m
    s2 = (gyg * xh).reshape(N, G, -1).sum(axis=2) / m
    r = np.repeat(rstd, C // G, axis=1)[:, :, None, None]
    grad_x = r * (gyg - xh * np.repeat(s2, C // G, axis=1)[:, :, None, None] - np.repeat(s1, C // G, axis=1)[:, :, None, None])
    return dict(grad_x=grad_x, grad_gamma=grad_gamma, grad_beta=grad_beta)
