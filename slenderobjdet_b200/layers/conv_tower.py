"""The dense-head towers -- ``nn.Conv2d(3x3, bias=False) -> nn.GroupNorm(32, C) -> nn.ReLU(inplace=True)`` stacks
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:644-675; run per FPN level :733-736;
fcos/fcos.py:494-538) -- on libslender_b200.

The plain convolution is the zero-offset specialisation of the deformable-convolution kernels: the same tcgen05
implicit GEMM (``sdb_dcn_forward_multi`` with ``offset == NULL``), whose gather loads ONE input row per (pixel, tap)
instead of four and skips the interpolation; grad_input is the same kernel run on grad_out with the transposed,
tap-reversed weights; grad_weight is the GEMM over the columns the forward saved.  One native call per pass covers
every FPN level of both towers.  bf16 tensor-core arithmetic only (rel <= 1e-2): bfloat16 tensors, bf16 autocast, or
``set_dcn_math("bf16")``; a float32 model that has not opted in is refused rather than silently demoted.
"""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

import sys

from .. import _lib
from .deform_conv import deform_conv_multi as _loaded  # noqa: F401  (the package re-exports a FUNCTION named deform_conv)
from .group_norm import GroupNormReLU, group_norm_relu_multi


_dc = sys.modules[__package__ + ".deform_conv"]   # the module, not the function of the same name


def _conv_plan(x, w, padding, dilation):
    g = _dc._geom(x, w, (1, 1), padding, dilation, 1, 1)
    cdt = _dc._compute_dtype(x)
    opted = cdt == torch.bfloat16 or _dc.get_dcn_math() == "bf16" or _dc._bf16_autocast()
    if not opted:
        raise RuntimeError("slender_b200: the tensor-core plain convolution computes with bf16 operands; use bfloat16 "
                           "tensors, torch.autocast(dtype=torch.bfloat16) or set_dcn_math('bf16') (float32 tensors are "
                           "not silently demoted)")
    iod = _lib.SDB_F32 if cdt == torch.float32 else _lib.SDB_BF16
    return g, cdt, iod, _lib.SDB_MATH_BF16


class _ConvMulti(Function):
    """Every 'same' convolution of a tower layer in one native call per pass.  Tensor arguments arrive flattened as
    inputs[n] + weights[k] + biases[k or 0]."""

    @staticmethod
    def forward(ctx, meta, *tensors):
        n, k, has_bias, wids, padding, dilation = meta
        xs, ws = list(tensors[:n]), list(tensors[n:n + k])
        bs = list(tensors[n + k:n + 2 * k]) if has_bias else [None] * k
        for t in xs:
            if t.dim() != 4:
                raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(t.dim()))
            if not t.is_cuda:
                raise NotImplementedError("slender_b200 convolution is not supported on CPUs!")
        w0 = ws[0]
        for w in ws:
            if tuple(w.shape) != tuple(w0.shape) or w.dtype != w0.dtype:
                raise RuntimeError("the convolutions of one multi call must share their weight shape and dtype")
        for x in xs:
            if x.shape[1] != w0.shape[1]:
                raise RuntimeError("invalid number of input planes, expected: %d, but got: %d" % (w0.shape[1], x.shape[1]))
        g, cdt, iod, mth = _conv_plan(xs[0], w0, padding, dilation)
        lib = _lib.lib()
        for x in xs:
            gi = _dc._geom(x, w0, (1, 1), padding, dilation, 1, 1)
            if not lib.sdb_dcn_supported(ctypes.byref(gi), iod, mth):
                raise RuntimeError("slender_b200: " + lib.sdb_last_error().decode())
        cx = [_dc._as(t, cdt) for t in xs]
        cw = [_dc._as(t, cdt) for t in ws]
        cb = [_dc._as(t, cdt) for t in bs]
        save = [bool(ctx.needs_input_grad[1 + n + wids[i]]) for i in range(n)]
        outs, packed = _dc._multi_forward(cx, [None] * n, [None] * n, cw, cb, wids, [-1] * n, g, iod, mth, cdt, save)
        ctx.save_for_backward(*tensors)
        ctx.meta_, ctx.packed_, ctx.g_, ctx.plan_ = meta, packed, g, (cdt, iod, mth)   # the backward uses the forward's arithmetic
        return tuple(o if o.dtype == xs[i].dtype else o.to(xs[i].dtype) for i, o in enumerate(outs))

    @staticmethod
    @once_differentiable
    def backward(ctx, *grad_outputs):
        n, k, has_bias, wids, padding, dilation = ctx.meta_
        tensors = ctx.saved_tensors
        xs, ws = list(tensors[:n]), list(tensors[n:n + k])
        bs = list(tensors[n + k:n + 2 * k]) if has_bias else [None] * k
        g = ctx.g_
        cdt, iod, mth = ctx.plan_
        need = ctx.needs_input_grad[1:]
        need_x = [bool(need[i]) for i in range(n)]
        need_w = [bool(need[n + j]) for j in range(k)]
        need_b = [bool(need[n + k + j]) for j in range(k)] if has_bias else [False] * k
        gxs, _, _, gws, gbs = _dc._multi_backward(
            [_dc._as(t, cdt) for t in xs], [None] * n, [None] * n, [_dc._as(t, cdt) for t in ws],
            [_dc._as(t, cdt) for t in bs], wids, [-1] * n, [_dc._as(gy, cdt) for gy in grad_outputs], ctx.packed_, g, iod,
            mth, cdt, need_x, [False] * n, [False] * n, need_w, need_b)
        out = [None]
        out += [None if t is None else _dc._as(t, xs[i].dtype) for i, t in enumerate(gxs)]
        out += [None if t is None else _dc._as(t, ws[j].dtype) for j, t in enumerate(gws)]
        if has_bias:
            out += [None if t is None else _dc._as(t, bs[j].dtype) for j, t in enumerate(gbs)]
        return tuple(out)


def conv2d_multi(inputs, weights, biases=None, padding=1, dilation=1, weight_ids=None):
    """``[F.conv2d(x_i, weights[weight_ids[i]], biases[...], stride=1, padding, dilation)]`` for 'same' convolutions
    (2 * padding == dilation * (k - 1)) in ONE native call per pass: every FPN level of a tower layer, or of both
    towers with two weights.  C_in % 64 == 0, C_out % 16 == 0, <= 256 channels (for grad_input also C_out % 64 == 0)."""
    inputs, weights = list(inputs), list(weights)
    n, k = len(inputs), len(weights)
    if n == 0:
        return []
    if n > _lib.SDB_MAX_PROBLEMS or k > _lib.SDB_MAX_WEIGHTS:
        raise ValueError("at most %d problems and %d weights per call" % (_lib.SDB_MAX_PROBLEMS, _lib.SDB_MAX_WEIGHTS))
    wids = tuple(int(v) for v in weight_ids) if weight_ids is not None else tuple([0] * n)
    if len(wids) != n or any(v < 0 or v >= k for v in wids):
        raise ValueError("weight_ids must name a weight for every input")
    has_bias = biases is not None and any(b is not None for b in biases)
    if has_bias and any(b is None for b in biases):
        raise ValueError("either every convolution of a call has a bias or none has")
    meta = (n, k, has_bias, wids, _pair(padding), _pair(dilation))
    args = inputs + weights + (list(biases) if has_bias else [])
    return list(_ConvMulti.apply(meta, *args))


class TowerConv2d(nn.Conv2d):
    """``nn.Conv2d`` ('same', stride 1, groups 1) whose forward and backward run on the tcgen05 kernels; constructor,
    parameters and state-dict keys are ``nn.Conv2d``'s."""

    def forward(self, input):
        if self.stride != (1, 1) or self.groups != 1 or self.padding_mode != "zeros" or isinstance(self.padding, str):
            raise RuntimeError("TowerConv2d: stride 1, groups 1, zero padding only")
        return conv2d_multi([input], [self.weight], [self.bias] if self.bias is not None else None, self.padding,
                            self.dilation)[0]


def build_tower(in_channels, feat_channels, stacked_convs=3, num_groups=32):
    """The reference's ``cls_convs`` / ``reg_convs`` ModuleList (reppointsv2.py:644-675) on these kernels.  Positions
    and parameter names match the reference (``0.weight``, ``1.weight``, ``1.bias``, ``3.weight`` ...), so its checkpoints
    load unchanged; the ReLU is fused into the normalisation, its slot holds an Identity."""
    layers = nn.ModuleList()
    for i in range(stacked_convs):
        chn = in_channels if i == 0 else feat_channels
        layers.append(TowerConv2d(chn, feat_channels, kernel_size=3, stride=1, padding=1, bias=False))
        layers.append(GroupNormReLU(num_groups * feat_channels // 256, feat_channels))
        layers.append(nn.Identity())
    return layers


def towers_forward(towers, features):
    """Runs ``len(towers)`` towers (e.g. ``[cls_convs, reg_convs]``) over every FPN level of ``features`` layer by
    layer -- the loop of reppointsv2.py:728-736 turned inside out: per layer ONE convolution call and ONE
    normalisation call over levels x towers.  -> ``[[tower_0 level outputs], [tower_1 level outputs], ...]``."""
    nt, nl = len(towers), len(features)
    if nt * nl > _lib.SDB_MAX_PROBLEMS or nt > _lib.SDB_MAX_WEIGHTS:
        raise ValueError("at most %d (tower, level) pairs and %d towers per call" % (_lib.SDB_MAX_PROBLEMS, _lib.SDB_MAX_WEIGHTS))
    cur = [f for _ in range(nt) for f in features]            # tower-major
    ids = [t for t in range(nt) for _ in range(nl)]
    depth = len(towers[0]) // 3
    for d in range(depth):
        convs = [tw[3 * d] for tw in towers]
        norms = [tw[3 * d + 1] for tw in towers]
        cur = conv2d_multi(cur, [c.weight for c in convs], [c.bias for c in convs] if convs[0].bias is not None else None,
                           convs[0].padding, convs[0].dilation, ids)
        cur = group_norm_relu_multi(cur, [m.weight for m in norms], [m.bias for m in norms], norms[0].num_groups,
                                    norms[0].eps, ids, relu=norms[0].relu)
    return [cur[t * nl:(t + 1) * nl] for t in range(nt)]
