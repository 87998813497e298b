"""Operator API mirroring ``detectron2.layers`` / ``slender_det.layers`` for the dense-head hot path."""
from .deform_conv import (DeformConv, ModulatedDeformConv, deform_conv, modulated_deform_conv, deform_conv_multi,
                          set_dcn_math, get_dcn_math, dcn_math, invalidate_prepared_weights, set_dcn_save_columns)
from .df_conv import DFConv2d
from .losses import (sigmoid_focal_loss, sigmoid_focal_loss_jit, sigmoid_focal_loss_from_class_idx,
                     iou_loss, box_iou_loss, smooth_l1_loss, smooth_l1_loss_with_weight, giou_loss,
                     compute_centerness_targets, compute_slender_centerness_targets)
from .reppoints_offset import reppoints_dcn_offset, dcn_base_offset
from .group_norm import GroupNormReLU, group_norm_relu, group_norm_relu_multi
from .conv_tower import TowerConv2d, conv2d_multi, build_tower, towers_forward

__all__ = [k for k in globals().keys() if not k.startswith("_")]
