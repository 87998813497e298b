"""CPU restatement (TEST INFRASTRUCTURE ONLY - never imported by the product path) of the RepPoints DCN offset
construction: /root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:638-642, 742-744 (no
channel flip) and rpd.py:105-110, 624-635 ((x, y) -> (y, x) flip per point first).  numpy float32, one
rounding per operation in the reference's order."""
import numpy as np


def dcn_base_offset(num_points):
    ks = int(np.sqrt(num_points))
    pad = int((ks - 1) / 2)
    base = np.arange(-pad, pad + 1).astype(np.float64)
    y = np.repeat(base, ks)
    x = np.tile(base, ks)
    return np.stack([y, x], axis=1).reshape(-1)          # reppointsv2.py:638-641


def dcn_offset(pts, gradient_mul=0.1, flip_xy=False):
    pts = np.asarray(pts, np.float32)
    n, c, h, w = pts.shape
    a, b = np.float32(1 - gradient_mul), np.float32(gradient_mul)
    gm = (a * pts).astype(np.float32) + (b * pts).astype(np.float32)     # :742-743 (detach is the identity forward)
    if flip_xy:                                                           # rpd.py:628-634
        gm = gm.reshape(n, c // 2, 2, h, w)[:, :, ::-1].reshape(n, c, h, w)
    base = dcn_base_offset(c // 2).astype(np.float32).reshape(1, -1, 1, 1)
    return (gm - base).astype(np.float32)                                 # :744


def dcn_offset_grad(grad_out, gradient_mul=0.1, flip_xy=False):
    g = np.asarray(grad_out, np.float32)
    n, c, h, w = g.shape
    out = (np.float32(gradient_mul) * g).astype(np.float32)
    if flip_xy:
        out = out.reshape(n, c // 2, 2, h, w)[:, :, ::-1].reshape(n, c, h, w)
    return out
