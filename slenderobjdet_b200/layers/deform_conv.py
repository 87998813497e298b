"""Drop-in replacement for ``detectron2.layers.deform_conv`` backed by libslender_b200.so.

Mirrors /root/reference/detectron2/detectron2/layers/deform_conv.py: same names
(``DeformConv``, ``ModulatedDeformConv``, ``deform_conv``, ``modulated_deform_conv``,
``_DeformConv``, ``_ModulatedDeformConv``), same argument order and meaning, same state-dict keys
(``weight`` [C_out, C_in/groups, kH, kW], ``bias``), same offset / mask layouts, same errors
(``ValueError`` for non-4D input :29-32 and too-small output :147-152, ``NotImplementedError`` on
CPU tensors :48-49, ``AssertionError`` when im2col_step does not divide N :52, ``RuntimeError`` for
shape violations, deform_conv_cuda.cu:140-270).

Differences, all internal: no ``columns`` / ``ones`` scratch tensors (the gather feeds the GEMM
directly) and ``im2col_step`` is accepted and validated but not used.

Arithmetic (``set_dcn_math`` / ``SDB_DCN_MATH``).  The reference computes float32 tensors in exact fp32
(im2col + fp32 GEMM, deform_conv_cuda_kernel.cu:486), so the default ``"auto"`` does the same: float32
tensors run the exact fp32 kernels (rel <= 1e-4 against the oracle) -- except that their FORWARD uses the
error-compensated ``tf32x3`` tensor-core kernel where the geometry allows (rel ~2e-5, still inside that bound;
``SDB_DCN_AUTO_TF32X3=0`` turns this off).  The tcgen05 tensor-core kernels (bf16
operands, fp32 accumulation in TMEM, rel <= 1e-2) are used when the TENSORS are bfloat16, under
``torch.autocast(dtype=torch.bfloat16)``, or when ``set_dcn_math("bf16")`` asks for them explicitly.
``set_dcn_math("tf32")`` / ``"tf32x3"`` run the FORWARD of float32 tensors on ``tcgen05.mma.kind::tf32`` (fp32
bilinear sampling; one pass with tf32-rounded operands, rel ~3e-4, or three error-compensated passes, rel ~2e-5 at
K = 2304); their backward stays on the exact fp32 kernels, so gradients keep fp32 accuracy.
float16 tensors are computed in float32 (the reference dispatches half too); float64 is refused.
"""
import contextlib
import ctypes
import math
import os
from functools import lru_cache

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _lib

_MATH = os.environ.get("SDB_DCN_MATH", "auto")  # "auto" | "bf16" | "fp32" | "tf32" | "tf32x3"
_MODES = ("auto", "bf16", "fp32", "tf32", "tf32x3")
_AUTO_TF32X3 = os.environ.get("SDB_DCN_AUTO_TF32X3", "1") != "0"   # 'auto' + float32 tensors: tf32x3 forward when supported


def set_dcn_math(mode):
    """'auto': follow the tensors -- float32 tensors use fp32-accurate kernels (forward: tf32x3 tensor cores where the
    geometry allows, else the exact fp32 kernel; backward: the exact fp32 kernels), bfloat16 tensors (or float32 under
    bf16 autocast) the tcgen05 tensor-core kernels when the geometry allows, else fp32;
    'bf16': require the tensor-core path whatever the tensor dtype (tolerance rel <= 1e-2);
    'fp32': always the exact fp32 path (rel <= 1e-4);
    'tf32' / 'tf32x3': float32 tensors, forward on kind::tf32 tensor cores in one pass (rel ~3e-4) / three
    error-compensated passes (rel ~2e-5, inside the 1e-4 bound of fp32 math), backward on the exact fp32 kernels; geometries the tensor-core kernel does not
    cover raise."""
    global _MATH
    assert mode in _MODES
    _MATH = mode


def get_dcn_math():
    return _MATH


@contextlib.contextmanager
def dcn_math(mode):
    old = _MATH
    set_dcn_math(mode)
    try:
        yield
    finally:
        set_dcn_math(old)


class _NewEmptyTensorOp(Function):
    """detectron2/layers/wrappers.py:29-38"""

    @staticmethod
    def forward(ctx, x, new_shape):
        ctx.shape = x.shape
        return x.new_empty(new_shape)

    @staticmethod
    def backward(ctx, grad):
        return _NewEmptyTensorOp.apply(grad, ctx.shape), None


def _geom(input, weight, stride, padding, dilation, groups, deformable_groups):
    return _lib.Geom(input.shape[0], input.shape[1], input.shape[2], input.shape[3], weight.shape[0],
                     weight.shape[2], weight.shape[3], stride[0], stride[1], padding[0], padding[1],
                     dilation[0], dilation[1], groups, deformable_groups)


def _check_shapes(input, offset, mask, weight, bias, g, out_hw):
    """shape_check (deform_conv_cuda.cu:140-270, :824-860) on the tensors the C ABI cannot see."""
    if weight.dim() != 4:
        raise RuntimeError("4D weight tensor (nOutputPlane,nInputPlane,kH,kW) expected, but got: %s" % weight.dim())
    if weight.shape[1] * g.groups != input.shape[1]:
        raise RuntimeError("invalid number of input planes, expected: %d, but got: %d"
                           % (weight.shape[1] * g.groups, input.shape[1]))
    n_off = g.deformable_groups * 2 * g.kH * g.kW
    if offset.dim() != 4 or offset.shape[0] != input.shape[0]:
        raise RuntimeError("invalid batch size of offset")
    if offset.shape[1] != n_off:
        raise RuntimeError("invalid number of channels of offset")
    if tuple(offset.shape[2:]) != tuple(out_hw):
        raise RuntimeError("invalid spatial size of offset, expected height: %d width: %d, but got height: %d width: %d"
                           % (out_hw[0], out_hw[1], offset.shape[2], offset.shape[3]))
    if mask is not None:
        if tuple(mask.shape) != (input.shape[0], g.deformable_groups * g.kH * g.kW, out_hw[0], out_hw[1]):
            raise RuntimeError("invalid shape of mask: expected %s, got %s"
                               % ((input.shape[0], g.deformable_groups * g.kH * g.kW) + tuple(out_hw), tuple(mask.shape)))
    if bias is not None and tuple(bias.shape) != (weight.shape[0],):
        raise RuntimeError("invalid shape of bias")
    for t in (weight, bias):
        if t is not None and t.dtype != input.dtype:
            raise RuntimeError("expected weight/bias dtype %s to match input dtype %s" % (t.dtype, input.dtype))
    for t in (offset, mask, weight, bias):
        if t is not None and t.device != input.device:
            raise RuntimeError("all tensors must be on the same CUDA device")


def _bf16_autocast():
    try:
        return torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16
    except Exception:  # pragma: no cover
        return False


def _pick_math(g, iod, autocast=False):
    lib = _lib.lib()
    if _MATH == "fp32":
        return _lib.SDB_MATH_FP32
    if _MATH in ("tf32", "tf32x3"):
        mth = _lib.SDB_MATH_TF32 if _MATH == "tf32" else _lib.SDB_MATH_TF32X3
        if not lib.sdb_dcn_supported(ctypes.byref(g), _lib.SDB_F32, mth):
            raise RuntimeError("slender_b200: " + lib.sdb_last_error().decode())
        return mth
    if _MATH == "auto" and iod == _lib.SDB_F32 and not autocast:
        # a float32 model keeps the reference's fp32 numerics unless told otherwise: the forward runs on tensor cores in
        # the error-compensated tf32x3 mode where the geometry allows (rel ~2e-5, inside the 1e-4 bound of fp32 math;
        # 14x the SIMT kernel), the backward always on the exact fp32 kernels
        if _AUTO_TF32X3 and lib.sdb_dcn_supported(ctypes.byref(g), _lib.SDB_F32, _lib.SDB_MATH_TF32X3):
            return _lib.SDB_MATH_TF32X3
        return _lib.SDB_MATH_FP32
    ok = bool(lib.sdb_dcn_supported(ctypes.byref(g), iod, _lib.SDB_MATH_BF16))
    if ok:
        return _lib.SDB_MATH_BF16
    if _MATH == "bf16":
        raise RuntimeError("slender_b200: " + lib.sdb_last_error().decode())
    return _lib.SDB_MATH_FP32


def _compute_dtype(t):
    """float32 / bfloat16 run natively; float16 (reference: AT_DISPATCH_FLOATING_TYPES_AND_HALF) is computed in
    float32; float64 (which the reference computes in double) is refused rather than silently downcast."""
    if t.dtype == torch.float64:
        raise RuntimeError("slender_b200: float64 deformable convolution is not implemented (float32 / bfloat16 / "
                           "float16 tensors only); cast the inputs explicitly")
    return t.dtype if t.dtype in (torch.float32, torch.bfloat16) else torch.float32


_PLAN_CACHE = {}


def _plan(input, weight, g):
    """-> (compute dtype, io_dtype enum, math enum, (Ho, Wo), workspace bytes per op, packed-input bytes).
    bf16 tensors whose geometry the tensor-core path does not cover are computed in float32 on the SIMT
    path.  Cached per (geometry, dtype, math mode): the public-API step is host-bound, and these are six
    ctypes round trips per call otherwise."""
    autocast = _bf16_autocast()
    key = (tuple(getattr(g, f) for f, _ in g._fields_), input.dtype, _MATH, autocast)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        return hit
    lib = _lib.lib()
    cdt = _compute_dtype(input)
    iod = _lib.SDB_F32 if cdt == torch.float32 else _lib.SDB_BF16
    mth = _pick_math(g, iod, autocast)
    if mth != _lib.SDB_MATH_BF16 and iod != _lib.SDB_F32:
        cdt, iod = torch.float32, _lib.SDB_F32
    ho, wo = ctypes.c_int32(0), ctypes.c_int32(0)
    _lib.check(lib.sdb_dcn_output_size(ctypes.byref(g), ho, wo))
    wsb = tuple(int(lib.sdb_dcn_workspace_bytes(op, ctypes.byref(g), iod, mth))
                for op in (_lib.SDB_OP_FORWARD, _lib.SDB_OP_BACKWARD_DATA, _lib.SDB_OP_BACKWARD_WEIGHT))
    pkb = int(lib.sdb_dcn_packed_input_bytes(ctypes.byref(g), mth))
    plan = (cdt, iod, mth, (ho.value, wo.value), wsb, pkb)
    if len(_PLAN_CACHE) < 4096:
        _PLAN_CACHE[key] = plan
    return plan


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def _as(t, dtype):
    """t as a contiguous tensor of `dtype` without touching it when it already is one"""
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


class _on_device(object):
    """torch.cuda.device(dev) only when dev is not already current (the context manager costs ~5 us)"""

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            return self.ctx.__exit__(*a)
        return False


# ---- prepared weights ---------------------------------------------------------------------------------------------
# The tensor-core kernels read the weights as pre-swizzled operand images (sdb_dcn_prepare_weights).  They depend on
# the weight VALUES only, so they are built once per weight version -- every forward / backward of every FPN level
# until the next optimiser step reuses them -- instead of once per native call.
_PREPARED = {}


def invalidate_prepared_weights():
    """Drop the cached operand images.  Only needed after mutating a weight through ``.data`` IN PLACE (which does
    not bump the tensor's version counter); optimiser steps, ``copy_`` / ``load_state_dict`` and ``.data = ...`` are
    detected automatically."""
    _PREPARED.clear()


def _prepared_weights(weight, bias, g, iod, mth, training=False):
    """-> uint8 tensor with the operand images of (weight, bias), cached on (tensor identity, version); None (= the
    native call builds what it needs itself, on the library's side stream beside its layout packs) for fp32 math and for
    training-mode bf16 calls, whose weights change every step so that a cached image would be rebuilt every step anyway."""
    if mth == _lib.SDB_MATH_FP32 or (training and mth == _lib.SDB_MATH_BF16):
        return None
    key = (id(weight), weight.device.index, mth)
    ver = (weight._version, weight.data_ptr(), None if bias is None else (id(bias), bias._version, bias.data_ptr()),
           iod, g.C_in, g.C_out, g.kH, g.kW)
    hit = _PREPARED.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == ver:
        return hit[2]
    lib = _lib.lib()
    nbytes = int(lib.sdb_dcn_prepared_weight_bytes(ctypes.byref(g), iod, mth))
    buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=weight.device)
    _lib.check(lib.sdb_dcn_prepare_weights(_lib.ptr(weight), _lib.ptr(bias), ctypes.byref(g), iod, mth, _lib.ptr(buf),
                                           _lib.stream_ptr(weight.device)))
    if len(_PREPARED) > 64:
        for k in [k for k, v in _PREPARED.items() if v[0]() is None]:
            del _PREPARED[k]
    import weakref
    _PREPARED[key] = (weakref.ref(weight), ver, buf)
    return buf


def _problem(x, off, m, wid=0, group=-1, out=None, packed=None, gy=None, gx=None, goff=None, gmask=None, cols=None):
    return _lib.Problem(x.shape[0], x.shape[2], x.shape[3], wid, group, 0, _lib.addr(x), _lib.addr(off), _lib.addr(m),
                        _lib.addr(out), _lib.addr(packed), _lib.addr(gy), _lib.addr(gx), _lib.addr(goff), _lib.addr(gmask),
                        _lib.addr(cols))


# Saved columns.  When a weight gradient will be wanted, the tensor-core forward also writes its sampled columns
# (kH*kW*C_in bf16 per output pixel, the reference's `columns` buffer) and the backward streams them into the
# weight-gradient GEMM instead of sampling the input a second time: ~2x faster weight gradient for 4.6 KB per output
# pixel (at C_in = 256, 3x3) held between forward and backward.  set_dcn_save_columns(False) / SDB_DCN_SAVE_COLUMNS=0
# trades the speed back for the memory.
_SAVE_COLUMNS = os.environ.get("SDB_DCN_SAVE_COLUMNS", "1") != "0"


def set_dcn_save_columns(on):
    global _SAVE_COLUMNS
    _SAVE_COLUMNS = bool(on)


def _multi_forward(xs, offs, masks, weights, biases, wids, groups, g, iod, mth, cdt, save_cols=None):
    """One sdb_dcn_forward_multi call.  xs / offs / masks: per-problem tensors already in the compute dtype;
    weights / biases: per-weight tensors; save_cols: per-problem flags "save the sampled columns for the backward".
    -> (outputs, list of (packed input | None, saved columns | None))."""
    lib = _lib.lib()
    n, dev = len(xs), xs[0].device
    tc = mth == _lib.SDB_MATH_BF16
    tf = mth in _lib.SDB_TF32_MODES
    gp = ctypes.byref(g)
    outs, packed, cols = [], [], []
    for i, x in enumerate(xs):
        gi = _lib.Geom(x.shape[0], g.C_in, x.shape[2], x.shape[3], g.C_out, g.kH, g.kW, g.sH, g.sW, g.pH, g.pW, g.dH,
                       g.dW, g.groups, g.deformable_groups)
        ho, wo = ctypes.c_int32(0), ctypes.c_int32(0)
        _lib.check(lib.sdb_dcn_output_size(ctypes.byref(gi), ho, wo))
        outs.append(torch.empty((x.shape[0], g.C_out, ho.value, wo.value), dtype=cdt, device=dev))
        packed.append(_ws(lib.sdb_dcn_packed_input_bytes(ctypes.byref(gi), mth), dev) if tc else None)
        want = tc and _SAVE_COLUMNS and save_cols is not None and save_cols[i] and x.shape[0] > 0
        cols.append(_ws(lib.sdb_dcn_columns_bytes(ctypes.byref(gi), mth), dev) if want else None)
    probs = (_lib.Problem * n)(*[_problem(xs[i], offs[i], masks[i], wids[i], groups[i], outs[i], packed[i], cols=cols[i])
                                 for i in range(n)])
    training = save_cols is not None and any(save_cols)   # a weight gradient will be wanted: training step
    prep = [_prepared_weights(w, b, g, iod, mth, training) for w, b in zip(weights, biases)]
    wts = (_lib.Weights * len(weights))(*[_lib.Weights(_lib.addr(w), _lib.addr(b), _lib.addr(p), None, None)
                                          for w, b, p in zip(weights, biases, prep)])
    wsb = int(lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, len(weights), gp, iod, mth, 0)) if (tc or tf) else 0
    ws = _ws(wsb, dev)
    with _on_device(dev):
        _lib.check(lib.sdb_dcn_forward_multi(probs, n, wts, len(weights), gp, iod, mth, _lib.ptr(ws), wsb,
                                             _lib.stream_ptr(dev)))
    return outs, list(zip(packed, cols))


def _multi_backward(xs, offs, masks, weights, biases, wids, groups, gys, packed, g, iod, mth, cdt, need_x, need_off,
                    need_mask, need_w, need_b, scale=1.0):
    """One sdb_dcn_backward_multi call: grad_offset / grad_mask, grad_input and grad_weight / grad_bias of every
    problem from one packed copy of grad_out.  need_*: per-problem (x, offset, mask) / per-weight (w, b) flags.
    -> (grad_x list, grad_offset list, grad_mask list, grad_weight list (fp32), grad_bias list (fp32))."""
    lib = _lib.lib()
    n, dev = len(xs), xs[0].device
    if mth in _lib.SDB_TF32_MODES:   # the tf32 modes are forward modes: gradients come from the exact fp32 kernels
        mth, packed = _lib.SDB_MATH_FP32, None
    tc = mth == _lib.SDB_MATH_BF16
    gp = ctypes.byref(g)
    # v1 computes grad_input and grad_offset together if either is needed (deform_conv.py:88); the kernels can skip either
    gxs = [torch.empty_like(xs[i]) if need_x[i] else None for i in range(n)]
    gos = [torch.empty_like(offs[i]) if need_off[i] else None for i in range(n)]
    gms = [torch.empty_like(masks[i]) if (masks[i] is not None and need_mask[i]) else None for i in range(n)]
    gws = [torch.zeros(w.shape, dtype=torch.float32, device=dev) if need_w[k] else None for k, w in enumerate(weights)]
    gbs = [torch.zeros((g.C_out,), dtype=torch.float32, device=dev) if (need_b[k] and biases[k] is not None) else None
           for k in range(len(weights))]
    saved = packed if packed is not None else [(None, None)] * n   # (packed input, saved columns) per problem
    probs = (_lib.Problem * n)(*[_problem(xs[i], offs[i], masks[i], wids[i], groups[i], None, saved[i][0], gys[i], gxs[i],
                                          gos[i], gms[i], cols=saved[i][1])
                                 for i in range(n)])
    # the weight VALUES are read only by grad_input / grad_offset / grad_mask; a grad_weight-only call needs no image
    reads = [any(wids[i] == k and (need_x[i] or need_off[i] or need_mask[i]) for i in range(n)) for k in range(len(weights))]
    prep = [_prepared_weights(w, b, g, iod, mth, any(need_w)) if r else None for w, b, r in zip(weights, biases, reads)]
    wts = (_lib.Weights * len(weights))(*[_lib.Weights(_lib.addr(w), _lib.addr(b), _lib.addr(p), _lib.addr(gw), _lib.addr(gb))
                                          for w, b, p, gw, gb in zip(weights, biases, prep, gws, gbs)])
    wsb = int(lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, len(weights), gp, iod, mth, 1)) if tc else 0
    ws = _ws(wsb, dev)
    with _on_device(dev):
        _lib.check(lib.sdb_dcn_backward_multi(probs, n, wts, len(weights), gp, iod, mth, float(scale), 0, _lib.ptr(ws), wsb,
                                              _lib.stream_ptr(dev)))
    return gxs, gos, gms, gws, gbs


def _forward_impl(input, offset, mask, weight, bias, g, save_cols=False):
    cdt, iod, mth, _, _, _ = _plan(input, weight, g)
    x, w, b = _as(input, cdt), _as(weight, cdt), _as(bias, cdt)
    off, m = _as(offset, torch.float32), _as(mask, torch.float32)
    outs, packed = _multi_forward([x], [off], [m], [w if w is not weight else weight], [b], [0], [-1], g, iod, mth, cdt,
                                  [save_cols])
    out = outs[0]
    return (out if out.dtype == input.dtype else out.to(input.dtype)), packed[0]


def _backward_impl(input, offset, mask, weight, grad_output, g, packed, need_data, need_weight, with_bias, scale=1.0,
                   bias=None):
    """-> grad_input, grad_offset, grad_mask, grad_weight, grad_bias (None where not requested)."""
    cdt, iod, mth, _, _, _ = _plan(input, weight, g)
    x, w, gy = _as(input, cdt), _as(weight, cdt), _as(grad_output, cdt)
    off, m = _as(offset, torch.float32), _as(mask, torch.float32)
    if with_bias and bias is None:   # only "is there a bias" matters to the backward
        bias = torch.zeros(g.C_out, dtype=cdt, device=input.device)
    b = _as(bias, cdt) if with_bias else None
    gxs, gos, gms, gws, gbs = _multi_backward([x], [off], [m], [w], [b], [0], [-1], [gy],
                                              [packed] if packed is not None else None, g, iod, mth, cdt,
                                              [need_data], [need_data], [need_data], [need_weight], [need_weight and with_bias],
                                              scale)
    gi, go, gm, gw, gb = gxs[0], gos[0], gms[0], gws[0], gbs[0]
    gi = _as(gi, input.dtype) if gi is not None else None
    go = _as(go, offset.dtype) if go is not None else None
    gm = _as(gm, mask.dtype) if gm is not None else None
    gw = _as(gw, weight.dtype) if gw is not None else None
    gb = _as(gb, weight.dtype) if gb is not None else None
    return gi, go, gm, gw, gb


class _DeformConv(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1, im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        ctx.stride = _pair(stride)
        ctx.padding = _pair(padding)
        ctx.dilation = _pair(dilation)
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.im2col_step = im2col_step

        out_size = _DeformConv._output_size(input, weight, ctx.padding, ctx.dilation, ctx.stride)
        if not input.is_cuda:
            raise NotImplementedError("Deformable Conv is not supported on CPUs!")
        cur_im2col_step = _DeformConv._cal_im2col_step(input.shape[0], ctx.im2col_step)
        assert (input.shape[0] % cur_im2col_step) == 0, "im2col step must divide batchsize"

        g = _geom(input, weight, ctx.stride, ctx.padding, ctx.dilation, groups, deformable_groups)
        _check_shapes(input, offset, None, weight, None, g, out_size[2:])
        output, packed = _forward_impl(input, offset, None, weight, None, g, ctx.needs_input_grad[2])
        ctx.save_for_backward(input, offset, weight)
        ctx.packed_ = packed  # (NHWC-bf16 copy of the input, saved columns), reused by backward (tensor-core path)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        if not grad_output.is_cuda:
            raise NotImplementedError("Deformable Conv is not supported on CPUs!")
        g = _geom(input, weight, ctx.stride, ctx.padding, ctx.dilation, ctx.groups, ctx.deformable_groups)
        need_data = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gi, go, _, gw, _ = _backward_impl(input, offset, None, weight, grad_output, g, ctx.packed_,
                                          need_data, ctx.needs_input_grad[2], False)
        return gi, go, gw, None, None, None, None, None, None

    @staticmethod
    def _output_size(input, weight, padding, dilation, stride):
        channels = weight.size(0)
        output_size = (input.size(0), channels)
        for d in range(input.dim() - 2):
            in_size = input.size(d + 2)
            kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
            output_size += ((in_size + (2 * padding[d]) - kernel) // stride[d] + 1,)
        if not all(map(lambda s: s > 0, output_size)):
            raise ValueError("convolution input is too small (output would be {})".format(
                "x".join(map(str, output_size))))
        return output_size

    @staticmethod
    @lru_cache(maxsize=128)
    def _cal_im2col_step(input_size, default_size):
        """Largest divisor of input_size that is <= default_size (deform_conv.py:155-176).  Kept for
        API compatibility; the fused kernels have no column buffer to chunk."""
        if input_size <= default_size:
            return input_size
        best_step = 1
        for step in range(2, min(int(math.sqrt(input_size)) + 1, default_size)):
            if input_size % step == 0:
                if input_size // step <= default_size:
                    return input_size // step
                best_step = step
        return best_step


class _ModulatedDeformConv(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1):
        ctx.stride = stride
        ctx.padding = padding
        ctx.dilation = dilation
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.with_bias = bias is not None
        if not input.is_cuda:
            raise NotImplementedError("Deformable Conv is not supported on CPUs!")
        if not input.is_contiguous() or not weight.is_contiguous():  # deform_conv_cuda.cu:824-825
            raise RuntimeError("input tensor has to be contiguous" if not input.is_contiguous()
                               else "weight tensor has to be contiguous")
        g = _geom(input, weight, _pair(stride), _pair(padding), _pair(dilation), groups, deformable_groups)
        out_shape = _ModulatedDeformConv._infer_shape(ctx, input, weight)
        if min(out_shape[2:]) <= 0:
            raise RuntimeError("convolution input is too small (output would be %s)" % (out_shape,))
        _check_shapes(input, offset, mask, weight, bias, g, out_shape[2:])
        output, packed = _forward_impl(input, offset, mask, weight, bias, g, ctx.needs_input_grad[3])
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(input, offset, mask, weight)
            ctx.packed_ = packed
            ctx.bias_ = bias.detach() if bias is not None else None
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError("Deformable Conv is not supported on CPUs!")
        input, offset, mask, weight = ctx.saved_tensors
        g = _geom(input, weight, _pair(ctx.stride), _pair(ctx.padding), _pair(ctx.dilation), ctx.groups,
                  ctx.deformable_groups)
        gi, go, gm, gw, gb = _backward_impl(input, offset, mask, weight, grad_output, g, ctx.packed_,
                                            True, True, ctx.with_bias, bias=ctx.bias_)
        return gi, go, gm, gw, gb, None, None, None, None, None

    @staticmethod
    def _infer_shape(ctx, input, weight):
        n = input.size(0)
        channels_out = weight.size(0)
        height, width = input.shape[2:4]
        kernel_h, kernel_w = weight.shape[2:4]
        height_out = (height + 2 * ctx.padding - (ctx.dilation * (kernel_h - 1) + 1)) // ctx.stride + 1
        width_out = (width + 2 * ctx.padding - (ctx.dilation * (kernel_w - 1) + 1)) // ctx.stride + 1
        return n, channels_out, height_out, width_out


class _DeformConvMulti(Function):
    """All deformable convolutions of a dense head in one native call per pass (sdb_dcn_forward_multi /
    sdb_dcn_backward_multi): every FPN level x every convolution.  Tensor arguments arrive flattened as
    inputs[n] + offsets[n] + masks[n or 0] + weights[k] + biases[k or 0]; ``meta`` describes the split."""

    @staticmethod
    def forward(ctx, meta, *tensors):
        n, k, has_mask, has_bias, wids, groups, stride, padding, dilation = meta
        groups = list(groups)
        xs = list(tensors[:n])
        offs = list(tensors[n:2 * n])
        pos = 2 * n
        masks = list(tensors[pos:pos + n]) if has_mask else [None] * n
        pos += n if has_mask else 0
        ws = list(tensors[pos:pos + k])
        pos += k
        bs = list(tensors[pos:pos + k]) if has_bias else [None] * k
        for t in xs:
            if t.dim() != 4:
                raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(t.dim()))
            if not t.is_cuda:
                raise NotImplementedError("Deformable Conv is not supported on CPUs!")
        w0 = ws[0]
        for w in ws:
            if tuple(w.shape) != tuple(w0.shape) or w.dtype != w0.dtype:
                raise RuntimeError("the convolutions of one multi call must share their weight shape and dtype")
        g = _geom(xs[0], w0, stride, padding, dilation, 1, 1)
        cdt, iod, mth, _, _, _ = _plan(xs[0], w0, g)
        for i in range(n):
            gi = _geom(xs[i], w0, stride, padding, dilation, 1, 1)
            ho, wo = ctypes.c_int32(0), ctypes.c_int32(0)
            try:
                _lib.check(_lib.lib().sdb_dcn_output_size(ctypes.byref(gi), ho, wo))
            except RuntimeError as e:
                raise ValueError(str(e))
            _check_shapes(xs[i], offs[i], masks[i], ws[wids[i]], bs[wids[i]], gi, (ho.value, wo.value))
        cx = [_as(t, cdt) for t in xs]
        co = [_as(t, torch.float32) for t in offs]
        cm = [_as(t, torch.float32) for t in masks]
        cw = [_as(t, cdt) for t in ws]
        cb = [_as(t, cdt) for t in bs]
        wpos = 2 * n + (n if has_mask else 0)   # weights' position among the tensor arguments (meta is argument 0)
        save = [bool(ctx.needs_input_grad[1 + wpos + wids[i]]) for i in range(n)]
        outs, packed = _multi_forward(cx, co, cm, cw, cb, wids, groups, g, iod, mth, cdt, save)
        ctx.save_for_backward(*tensors)
        ctx.meta_, ctx.groups_, ctx.packed_, ctx.g_ = meta, groups, packed, g
        return tuple(o if o.dtype == xs[i].dtype else o.to(xs[i].dtype) for i, o in enumerate(outs))

    @staticmethod
    @once_differentiable
    def backward(ctx, *grad_outputs):
        n, k, has_mask, has_bias, wids, _, stride, padding, dilation = ctx.meta_
        tensors = ctx.saved_tensors
        xs, offs = list(tensors[:n]), list(tensors[n:2 * n])
        pos = 2 * n
        masks = list(tensors[pos:pos + n]) if has_mask else [None] * n
        pos += n if has_mask else 0
        ws = list(tensors[pos:pos + k])
        pos += k
        bs = list(tensors[pos:pos + k]) if has_bias else [None] * k
        g = ctx.g_
        cdt, iod, mth, _, _, _ = _plan(xs[0], ws[0], g)
        need = ctx.needs_input_grad[1:]
        need_x = [bool(need[i]) for i in range(n)]
        need_off = [bool(need[n + i]) for i in range(n)]
        pos = 2 * n
        need_mask = [bool(need[pos + i]) for i in range(n)] if has_mask else [False] * n
        pos += n if has_mask else 0
        need_w = [bool(need[pos + j]) for j in range(k)]
        pos += k
        need_b = [bool(need[pos + j]) for j in range(k)] if has_bias else [False] * k
        gys = [_as(gy, cdt) for gy in grad_outputs]
        gxs, gos, gms, gws, gbs = _multi_backward(
            [_as(t, cdt) for t in xs], [_as(t, torch.float32) for t in offs], [_as(t, torch.float32) for t in masks],
            [_as(t, cdt) for t in ws], [_as(t, cdt) for t in bs], wids, ctx.groups_, gys, ctx.packed_, g, iod, mth, cdt,
            need_x, need_off, need_mask, need_w, need_b)
        out = [None]
        out += [None if t is None else _as(t, xs[i].dtype) for i, t in enumerate(gxs)]
        out += [None if t is None else _as(t, offs[i].dtype) for i, t in enumerate(gos)]
        if has_mask:
            out += [None if t is None else _as(t, masks[i].dtype) for i, t in enumerate(gms)]
        out += [None if t is None else _as(t, ws[j].dtype) for j, t in enumerate(gws)]
        if has_bias:
            out += [None if t is None else _as(t, bs[j].dtype) for j, t in enumerate(gbs)]
        return tuple(out)


def deform_conv_multi(inputs, offsets, weights, stride=1, padding=0, dilation=1, masks=None, biases=None,
                      weight_ids=None):
    """Every deformable convolution of a dense head in ONE native call per pass.

    The reference loops ``deform_conv`` over FPN levels and branches (reppointsv2.py:728-752: 10 forward and 20
    backward native calls per step for the RepPoints head).  Here ``inputs`` / ``offsets`` (and ``masks`` for DCNv2)
    list one entry per (level, branch) problem, ``weights`` is one tensor or a list of up to four that share their
    shape (e.g. ``[cls_conv.weight, refine_conv.weight]``) and ``weight_ids[i]`` says which of them problem ``i``
    uses.  Problems that are given the SAME offset tensor object (the two DCNs of a RepPoints level consume one
    ``dcn_offset``) share the transposed sampling index in the backward.  Semantics per problem are exactly those of
    ``deform_conv`` / ``modulated_deform_conv``; groups = deformable_groups = 1.  Returns the list of outputs."""
    single_w = torch.is_tensor(weights)
    ws = [weights] if single_w else list(weights)
    n = len(inputs)
    if n == 0:
        return []
    if len(offsets) != n or (masks is not None and len(masks) != n):
        raise ValueError("inputs, offsets and masks must have one entry per problem")
    if n > _lib.SDB_MAX_PROBLEMS or len(ws) > _lib.SDB_MAX_WEIGHTS:
        raise ValueError("at most %d problems and %d weight tensors per call" % (_lib.SDB_MAX_PROBLEMS, _lib.SDB_MAX_WEIGHTS))
    wids = [0] * n if weight_ids is None else [int(v) for v in weight_ids]
    if biases is not None and torch.is_tensor(biases):
        biases = [biases]
    if biases is not None and any(b is None for b in biases):
        raise ValueError("biases: give one tensor per weight, or None for no bias at all")
    # offsets that are the same tensor object (with the same mask and input size) form one offset group: the backward
    # builds their transposed sampling index once
    groups, seen = [], {}
    for i in range(n):
        key = (id(offsets[i]), None if masks is None else id(masks[i]), tuple(inputs[i].shape))
        groups.append(seen.setdefault(key, len(seen)))
    meta = (n, len(ws), masks is not None, biases is not None, tuple(wids), tuple(groups), _pair(stride), _pair(padding),
            _pair(dilation))
    tensors = list(inputs) + list(offsets) + (list(masks) if masks is not None else []) + ws + \
        (list(biases) if biases is not None else [])
    return list(_DeformConvMulti.apply(meta, *tensors))


deform_conv = _DeformConv.apply
modulated_deform_conv = _ModulatedDeformConv.apply


def _empty_output(x, weight, padding, dilation, kernel_size, stride):
    output_shape = [(i + 2 * p - (di * (k - 1) + 1)) // s + 1
                    for i, p, di, k, s in zip(x.shape[-2:], padding, dilation, kernel_size, stride)]
    return _NewEmptyTensorOp.apply(x, [x.shape[0], weight.shape[0]] + output_shape)


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=False, norm=None, activation=None):
        """Deformable convolution (DCN v1).  Arguments as detectron2.layers.DeformConv
        (deform_conv.py:309-359): ``deformable_groups``, optional ``norm`` module and
        ``activation`` callable applied after the convolution."""
        super(DeformConv, self).__init__()
        assert not bias
        assert in_channels % groups == 0, "in_channels {} cannot be divisible by groups {}".format(
            in_channels, groups)
        assert out_channels % groups == 0, "out_channels {} cannot be divisible by groups {}".format(
            out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.norm = norm
        self.activation = activation
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.bias = None
        nn.init.kaiming_uniform_(self.weight, nonlinearity="relu")

    def forward(self, x, offset):
        if x.numel() == 0:
            return _empty_output(x, self.weight, self.padding, self.dilation, self.kernel_size, self.stride)
        x = deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                        self.deformable_groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x

    def forward_multi(self, xs, offsets):
        """``[self(x, offset) for x, offset in zip(xs, offsets)]`` for all FPN levels in one native call per pass
        (``deform_conv_multi``); norm / activation are applied per level as in ``forward``."""
        if self.groups != 1 or self.deformable_groups != 1 or any(x.numel() == 0 for x in xs):
            return [self(x, o) for x, o in zip(xs, offsets)]
        ys = deform_conv_multi(xs, offsets, self.weight, self.stride, self.padding, self.dilation)
        if self.norm is not None:
            ys = [self.norm(y) for y in ys]
        if self.activation is not None:
            ys = [self.activation(y) for y in ys]
        return ys

    def extra_repr(self):
        tmpstr = "in_channels=" + str(self.in_channels)
        tmpstr += ", out_channels=" + str(self.out_channels)
        tmpstr += ", kernel_size=" + str(self.kernel_size)
        tmpstr += ", stride=" + str(self.stride)
        tmpstr += ", padding=" + str(self.padding)
        tmpstr += ", dilation=" + str(self.dilation)
        tmpstr += ", groups=" + str(self.groups)
        tmpstr += ", deformable_groups=" + str(self.deformable_groups)
        tmpstr += ", bias=False"
        return tmpstr


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=True, norm=None, activation=None):
        """Modulated deformable convolution (DCN v2).  Arguments as
        detectron2.layers.ModulatedDeformConv (deform_conv.py:406-452); scalar stride / padding /
        dilation as in the reference."""
        super(ModulatedDeformConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.with_bias = bias
        self.norm = norm
        self.activation = activation
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.bias = None
        nn.init.kaiming_uniform_(self.weight, nonlinearity="relu")
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)

    def forward(self, x, offset, mask):
        if x.numel() == 0:
            return _empty_output(x, self.weight, _pair(self.padding), _pair(self.dilation), self.kernel_size,
                                 _pair(self.stride))
        x = modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                  self.dilation, self.groups, self.deformable_groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x

    def extra_repr(self):
        tmpstr = "in_channels=" + str(self.in_channels)
        tmpstr += ", out_channels=" + str(self.out_channels)
        tmpstr += ", kernel_size=" + str(self.kernel_size)
        tmpstr += ", stride=" + str(self.stride)
        tmpstr += ", padding=" + str(self.padding)
        tmpstr += ", dilation=" + str(self.dilation)
        tmpstr += ", groups=" + str(self.groups)
        tmpstr += ", deformable_groups=" + str(self.deformable_groups)
        tmpstr += ", bias=" + str(self.with_bias)
        return tmpstr
