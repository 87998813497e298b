"""RepPoints DCN offset construction, fused (first piece).

Mirrors /root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:638-642, 742-744 and
rpd.py:105-110, 624-635::

    pts_out_init_grad_mul = (1 - gradient_mul) * pts_out_init.detach() + gradient_mul * pts_out_init
    dcn_offset = pts_out_init_grad_mul - dcn_base_offset        # rpd.py first flips (x, y) -> (y, x) per point

as one kernel forward (bit-identical: the reference's four float32 roundings are kept) and one backward
(``grad = gradient_mul * grad_out``), instead of four elementwise launches and three temporaries per level.
"""
import math

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib


def dcn_base_offset(num_points, device=None, dtype=torch.float32):
    """The reference's ``dcn_base_offset`` buffer: [1, 2*num_points, 1, 1], (y, x) per point, row-major grid."""
    ks = int(math.isqrt(num_points))
    assert ks * ks == num_points, "The points number should be a square number."
    pad = (ks - 1) // 2
    base = torch.arange(-pad, pad + 1, dtype=torch.float64)
    y = base.repeat_interleave(ks)
    x = base.repeat(ks)
    return torch.stack([y, x], dim=1).reshape(-1).to(dtype).view(1, -1, 1, 1).to(device)


class _RepPointsDcnOffset(Function):
    @staticmethod
    def forward(ctx, pts, gradient_mul, flip_xy):
        if not pts.is_cuda:
            raise NotImplementedError("slender_b200: CUDA tensors only (no CPU fallback)")
        if pts.dim() != 4 or pts.shape[1] % 2:
            raise ValueError("expected pts_out_init of shape [N, 2*num_points, H, W]")
        k = pts.shape[1] // 2
        ks = int(math.isqrt(k))
        if ks * ks != k:
            raise ValueError("The points number should be a square number.")
        p = pts.detach()
        p = p if p.dtype == torch.float32 else p.float()
        p = p if p.is_contiguous() else p.contiguous()
        out = torch.empty_like(p)
        N, _, H, W = p.shape
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().sdb_reppoints_dcn_offset(_lib.ptr(p), N, ks, H, W, float(gradient_mul), int(flip_xy),
                                                           _lib.ptr(out), _lib.stream_ptr(p.device)))
        ctx.cfg = (N, ks, H, W, float(gradient_mul), int(flip_xy), pts.dtype)
        return out.to(pts.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        N, ks, H, W, gm, flip, dt = ctx.cfg
        g = grad_out if grad_out.dtype == torch.float32 else grad_out.float()
        g = g if g.is_contiguous() else g.contiguous()
        gp = torch.empty_like(g)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().sdb_reppoints_dcn_offset_backward(_lib.ptr(g), N, ks, H, W, gm, flip, _lib.ptr(gp),
                                                                    _lib.stream_ptr(g.device)))
        return gp.to(dt), None, None


def reppoints_dcn_offset(pts_out_init, gradient_mul=0.1, flip_xy=False):
    """``dcn_offset`` for the RepPoints refine / classify DeformConvs from the initial point offsets.
    ``flip_xy=False`` is reppointsv2.py (no flip), ``True`` is rpd.py / rpd_centerness.py."""
    return _RepPointsDcnOffset.apply(pts_out_init, gradient_mul, flip_xy)
