"""slenderobjdet_b200 -- B200 (sm_100a) kernels for SlenderObjDet's dense-head hot path.

Deformable convolution v1/v2 forward + backward, IoU/top-k label assignment and the fused head
losses, behind the reference's ``detectron2.layers`` / ``slender_det.layers`` operator API.
Everything executes in libslender_b200.so (hand-written CUDA, C ABI in include/slender_b200.h);
there is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401
from .layers import (DeformConv, ModulatedDeformConv, DFConv2d, deform_conv, modulated_deform_conv,  # noqa: F401
                     deform_conv_multi, set_dcn_math, get_dcn_math, dcn_math, invalidate_prepared_weights,
                     set_dcn_save_columns)
from .matchers import Matcher, TopKMatcher, pairwise_iou  # noqa: F401

__version__ = "0.1.0"
