"""A ``detectron2._C``-compatible surface for the five deformable-convolution functions, backed by libslender_b200.so.

The reference binds its native code as ``detectron2._C`` (detectron2/detectron2/layers/csrc/vision.cpp:76-92) and calls
it from ``detectron2/layers/deform_conv.py`` with POSITIONAL arguments, caller-allocated outputs and pre-zeroed
gradients.  This module exposes the same five names with the same argument order and in-place behaviour, so the
reference's own ``deform_conv.py`` runs unmodified on these kernels with one line changed::

    from slenderobjdet_b200 import d2_C as _C        # instead of: from detectron2 import _C

Behaviour kept from the reference (file:line in deform_conv_cuda.cu unless noted):
  * v1 passes W before H (kW, kH, dW, dH, padW, padH, dilationW, dilationH), v2 H before W (deform_conv.h:116-375);
  * ``output`` / ``grad_offset`` / ``grad_mask`` are overwritten, ``grad_input`` / ``grad_weight`` / ``grad_bias`` are
    accumulated into (:770-777, :1102-1113; the Python side zeroes them first, deform_conv.py:89-90, :113, :242-246);
  * ``columns`` / ``ones`` are accepted and ignored (the reference reallocates them internally, :345-352; the fused
    kernels have no column buffer);
  * v1 makes its inputs contiguous internally (:312-314); v2 requires contiguous input and weight (:824-825);
  * shape violations raise ``RuntimeError`` (shape_check :140-270); v1 returns ``1`` like the reference.
Arithmetic follows ``set_dcn_math`` (slenderobjdet_b200.layers.deform_conv): float32 tensors use the exact fp32
kernels unless bf16 / tf32 tensor-core math was selected explicitly; bfloat16 tensors use the tcgen05 kernels.
"""
import ctypes
import importlib

from . import _lib

# the MODULE (the package re-exports a function of the same name)
_dc = importlib.import_module(".layers.deform_conv", __package__)


def _geom(input, weight, kH, kW, sH, sW, pH, pW, dH, dW, group, deformable_group):
    if input.dim() != 4:
        raise RuntimeError("4D input tensor expected but got: %s" % input.dim())
    if weight.dim() != 4:
        raise RuntimeError("4D weight tensor (nOutputPlane,nInputPlane,kH,kW) expected, but got: %s" % weight.dim())
    if (weight.shape[2], weight.shape[3]) != (kH, kW):
        raise RuntimeError("kernel size should be consistent with weight, but got kH: %d kW: %d weight.size(2): %d, "
                           "weight.size(3): %d" % (kH, kW, weight.shape[2], weight.shape[3]))
    return _lib.Geom(input.shape[0], input.shape[1], input.shape[2], input.shape[3], weight.shape[0], kH, kW,
                     sH, sW, pH, pW, dH, dW, group, deformable_group)


def _out_hw(g):
    ho, wo = ctypes.c_int32(0), ctypes.c_int32(0)
    _lib.check(_lib.lib().sdb_dcn_output_size(ctypes.byref(g), ho, wo))
    return ho.value, wo.value


def _store(dst, src, accumulate):
    """Write a result into the caller's tensor (any dtype / stride), overwriting or accumulating."""
    if dst is src:
        return
    if accumulate:
        dst.add_(src.to(dst.dtype))
    else:
        dst.copy_(src)


def deform_conv_forward(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH,
                        group, deformable_group, im2col_step):
    """vision.cpp:79 -> deform_conv_forward_cuda (deform_conv_cuda.cu:272-438)."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    g = _geom(input, weight, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
    if input.shape[0] % im2col_step != 0:
        raise RuntimeError("im2col step must divide batchsize")   # :338 (batchSize % im2col_step == 0)
    _dc._check_shapes(input, offset, None, weight, None, g, _out_hw(g))
    out, _ = _dc._forward_impl(input, offset, None, weight, None, g)
    if tuple(output.shape) != tuple(out.shape):
        output.resize_(out.shape)                                   # the reference views / resizes `output` (:354-363)
    output.copy_(out)
    return 1


def deform_conv_backward_input(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH, padW,
                               padH, dilationW, dilationH, group, deformable_group, im2col_step):
    """vision.cpp:80-83 -> deform_conv_backward_input_cuda (:440-628): grad_input accumulated (atomics in the reference),
    grad_offset stored."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    g = _geom(input, weight, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
    _dc._check_shapes(input, offset, None, weight, None, g, _out_hw(g))
    gi, go, _, _, _ = _dc._backward_impl(input, offset, None, weight, gradOutput, g, None, True, False, False)
    _store(gradInput, gi, True)
    _store(gradOffset, go, False)
    return 1


def deform_conv_backward_filter(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH,
                                dilationW, dilationH, group, deformable_group, scale, im2col_step):
    """vision.cpp:84-87 -> deform_conv_backward_parameters_cuda (:630-802): gradWeight += scale * dY . col^T."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    g = _geom(input, gradWeight, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
    _dc._check_shapes(input, offset, None, gradWeight, None, g, _out_hw(g))
    _, _, _, gw, _ = _dc._backward_impl(input, offset, None, gradWeight, gradOutput, g, None, False, True, False,
                                        scale=float(scale))
    _store(gradWeight, gw, True)
    return 1


def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h,
                                  stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    """vision.cpp:88-91 -> modulated_deform_conv_cuda_forward (:804-927)."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")    # :824
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")   # :825
    g = _geom(input, weight, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
              deformable_group)
    b = bias if with_bias else None
    _dc._check_shapes(input, offset, mask, weight, b, g, _out_hw(g))
    out, _ = _dc._forward_impl(input, offset, mask, weight, b, g)
    if tuple(output.shape) != tuple(out.shape):
        output.resize_(out.shape)                                   # :858-860 (output.view(...).zero_())
    output.copy_(out)


def modulated_deform_conv_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight, grad_bias,
                                   grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                   pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    """vision.cpp:92 -> modulated_deform_conv_cuda_backward (:929-1129): all five gradients in one call."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")
    g = _geom(input, weight, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
              deformable_group)
    _dc._check_shapes(input, offset, mask, weight, bias if with_bias else None, g, _out_hw(g))
    gi, go, gm, gw, gb = _dc._backward_impl(input, offset, mask, weight, grad_output, g, None, True, True, bool(with_bias),
                                            bias=bias if with_bias else None)
    _store(grad_input, gi, True)
    _store(grad_offset, go, False)
    _store(grad_mask, gm, False)
    _store(grad_weight, gw, True)
    if with_bias:
        _store(grad_bias, gb, True)
