"""GroupNorm + ReLU of the dense-head towers on libslender_b200 (csrc/gn.cu).

The reference's towers are ``nn.Conv2d(3x3, bias=False) -> nn.GroupNorm(32, C) -> nn.ReLU(inplace=True)`` stacks
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:644-675, run per FPN level :733-736).
``group_norm_relu_multi`` normalises every level (and both towers) of one layer in a single native call per pass;
``GroupNormReLU`` is the one-tensor module with ``nn.GroupNorm``'s constructor, parameters and state-dict keys.
There is no CPU or eager fallback: CPU tensors raise ``NotImplementedError`` like the deformable convolution does.
"""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib


def _call(fn, xs, ys, gys, gxs, stats, pids, gammas, betas, ggs, gbs, G, eps, relu):
    lib = _lib.lib()
    n, dev = len(xs), xs[0].device
    C = xs[0].shape[1]
    ten = (_lib.GnTensor * n)(*[
        _lib.GnTensor(_lib.addr(xs[i]), _lib.addr(ys[i]) if ys else None, _lib.addr(gys[i]) if gys else None,
                      _lib.addr(gxs[i]) if gxs else None, _lib.addr(stats[i]), xs[i].shape[0],
                      xs[i].shape[2] * xs[i].shape[3], pids[i], 0) for i in range(n)])
    k = len(gammas)
    par = (_lib.GnParams * k)(*[_lib.GnParams(_lib.addr(gammas[j]), _lib.addr(betas[j]), _lib.addr(ggs[j]) if ggs else None,
                                              _lib.addr(gbs[j]) if gbs else None) for j in range(k)])
    wsb = int(lib.sdb_gn_relu_workspace_bytes(ten, n, C, G))
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    _lib.check(fn(ten, n, par, k, C, G, float(eps), int(relu), _lib.io_dtype(xs[0]), _lib.ptr(ws), wsb, _lib.stream_ptr(dev)))


class _GroupNormReLUMulti(Function):
    @staticmethod
    def forward(ctx, meta, *tensors):
        n, k, pids, G, eps, relu = meta
        xs = [t.contiguous() for t in tensors[:n]]
        gammas = [t.detach().float().contiguous() for t in tensors[n:n + k]]
        betas = [t.detach().float().contiguous() for t in tensors[n + k:n + 2 * k]]
        for x in xs:
            if x.dim() != 4:
                raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(x.dim()))
            if not x.is_cuda:
                raise NotImplementedError("slender_b200 GroupNorm is not supported on CPUs!")
            if x.shape[1] != xs[0].shape[1] or x.dtype != xs[0].dtype:
                raise RuntimeError("the tensors of one call must share their channel count and dtype")
        C = xs[0].shape[1]
        if C % G != 0:
            raise ValueError("num_channels must be divisible by num_groups")
        for t in gammas + betas:
            if tuple(t.shape) != (C,):
                raise RuntimeError("weight / bias must have shape (%d,)" % C)
        ys = [torch.empty_like(x) for x in xs]
        stats = [torch.empty((x.shape[0], G, 2), dtype=torch.float32, device=x.device) for x in xs]
        with torch.cuda.device(xs[0].device):
            _call(_lib.lib().sdb_gn_relu_forward, xs, ys, None, None, stats, pids, gammas, betas, None, None, G, eps, relu)
        ctx.save_for_backward(*xs, *stats, *gammas, *betas)
        ctx.meta_ = meta
        ctx.param_dtypes_ = [t.dtype for t in tensors[n:n + 2 * k]]
        return tuple(ys)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        n, k, pids, G, eps, relu = ctx.meta_
        sv = ctx.saved_tensors
        xs, stats = list(sv[:n]), list(sv[n:2 * n])
        gammas, betas = list(sv[2 * n:2 * n + k]), list(sv[2 * n + k:2 * n + 2 * k])
        need = ctx.needs_input_grad[1:]
        gys = [g.to(xs[i].dtype).contiguous() for i, g in enumerate(grads)]
        gxs = [torch.empty_like(x) for x in xs]   # dx is one extra store of a pass that runs anyway
        ggs = [torch.zeros_like(g) for g in gammas]
        gbs = [torch.zeros_like(b) for b in betas]
        with torch.cuda.device(xs[0].device):
            _call(_lib.lib().sdb_gn_relu_backward, xs, None, gys, gxs, stats, pids, gammas, betas, ggs, gbs, G, eps, relu)
        out = [None]
        out += [gxs[i] if need[i] else None for i in range(n)]
        out += [ggs[j].to(ctx.param_dtypes_[j]) if need[n + j] else None for j in range(k)]
        out += [gbs[j].to(ctx.param_dtypes_[k + j]) if need[n + k + j] else None for j in range(k)]
        return tuple(out)


def group_norm_relu_multi(inputs, weights, biases, num_groups, eps=1e-5, param_ids=None, relu=True):
    """``[relu(group_norm(x_i, num_groups, weights[param_ids[i]], biases[param_ids[i]], eps))]`` for every tensor of
    ``inputs`` (NCHW, same channel count and dtype; typically the FPN levels of one tower layer, or of both towers with
    two parameter sets) in one native call per pass.  <= 16 tensors, <= 4 parameter sets."""
    inputs, weights, biases = list(inputs), list(weights), list(biases)
    n, k = len(inputs), len(weights)
    if len(biases) != k:
        raise ValueError("one bias per weight")
    pids = tuple(int(p) for p in param_ids) if param_ids is not None else tuple([0] * n)
    if len(pids) != n or any(p < 0 or p >= k for p in pids):
        raise ValueError("param_ids must name a weight for every input")
    if n == 0:
        return []
    meta = (n, k, pids, int(num_groups), float(eps), bool(relu))
    return list(_GroupNormReLUMulti.apply(meta, *inputs, *weights, *biases))


def group_norm_relu(input, num_groups, weight, bias, eps=1e-5, relu=True):
    """relu(F.group_norm(input, num_groups, weight, bias, eps)) in two launches"""
    return group_norm_relu_multi([input], [weight], [bias], num_groups, eps, None, relu)[0]


class GroupNormReLU(nn.GroupNorm):
    """``nn.GroupNorm(num_groups, num_channels)`` followed by ``nn.ReLU`` as one module: same constructor, parameters
    (``weight``, ``bias``) and state-dict keys as ``nn.GroupNorm``, so a reference checkpoint's
    ``cls_convs.1.weight`` / ``.bias`` load unchanged."""

    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True, relu=True):
        super().__init__(num_groups, num_channels, eps=eps, affine=affine)
        self.relu = relu

    def forward(self, input):
        w = self.weight if self.affine else torch.ones(self.num_channels, device=input.device)
        b = self.bias if self.affine else torch.zeros(self.num_channels, device=input.device)
        return group_norm_relu(input, self.num_groups, w, b, self.eps, self.relu)
