"""``DFConv2d`` -- offset(+mask)-predicting conv followed by DCN v1/v2.

Same constructor, attributes, parameter names (``offset.weight``, ``offset.bias``, ``conv.weight``,
``conv.bias``) and forward semantics as /root/reference/slender_det/layers/df_conv.py:6-79: the
plain ``offset`` convolution predicts ``deformable_groups * 2*k*k`` offset channels (v1) or
``3*k*k`` (v2: first 2/3 are offsets, last 1/3 mask logits passed through a sigmoid, :75-78).
The deformable convolution itself runs on libslender_b200's kernels.
"""
from torch import nn

from .deform_conv import DeformConv, ModulatedDeformConv


class DFConv2d(nn.Module):
    """Deformable convolution layer"""

    def __init__(self, in_channels, out_channels, with_modulated_dcn=True, kernel_size=3, stride=1,
                 groups=1, padding=1, dilation=1, deformable_groups=1, bias=False):
        super().__init__()
        if isinstance(kernel_size, (list, tuple)):
            assert len(kernel_size) == 2
            taps = kernel_size[0] * kernel_size[1]
        else:
            taps = kernel_size * kernel_size
        self.offset_base_channels = taps
        self.with_modulated_dcn = with_modulated_dcn
        per_tap = 3 if with_modulated_dcn else 2  # (dy, dx[, mask]) per tap
        # detectron2.layers.Conv2d without norm/activation is a plain nn.Conv2d (same state-dict keys)
        self.offset = nn.Conv2d(in_channels, deformable_groups * taps * per_tap, kernel_size=kernel_size,
                                stride=stride, padding=padding, groups=1, dilation=dilation)
        nn.init.kaiming_uniform_(self.offset.weight, a=1)
        nn.init.constant_(self.offset.bias, 0.0)
        dcn = ModulatedDeformConv if with_modulated_dcn else DeformConv
        self.conv = dcn(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                        dilation=dilation, groups=groups, deformable_groups=deformable_groups, bias=bias)
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation

    def forward(self, x):
        assert x.numel() > 0, "only non-empty tensors are supported"
        pred = self.offset(x)
        if not self.with_modulated_dcn:
            return self.conv(x, pred)
        split = self.offset_base_channels * 2
        return self.conv(x, pred[:, :split, :, :], pred[:, split:, :, :].sigmoid())
