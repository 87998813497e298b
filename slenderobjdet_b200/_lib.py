"""ctypes binding of libslender_b200.so (the C ABI declared in include/slender_b200.h).

The library is the product: there is NO Python / CPU fallback.  If the shared object is missing
(not built) every op raises ``RuntimeError`` telling the user how to build it.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# SDB_LIB_PATH: developer switch for A/B timing of two builds on the same GPU box (tools/ab_bench.sh)
LIB_PATH = os.environ.get("SDB_LIB_PATH") or os.path.join(_PKG, "libslender_b200.so")

SDB_F32, SDB_BF16 = 0, 1
SDB_MATH_FP32, SDB_MATH_BF16, SDB_MATH_TF32, SDB_MATH_TF32X3 = 0, 1, 2, 3
SDB_TF32_MODES = (SDB_MATH_TF32, SDB_MATH_TF32X3)
SDB_OP_FORWARD, SDB_OP_BACKWARD_DATA, SDB_OP_BACKWARD_WEIGHT = 0, 1, 2
SDB_LOSS_IOU, SDB_LOSS_LINEAR_IOU, SDB_LOSS_GIOU, SDB_LOSS_SMOOTH_L1, SDB_LOSS_GIOU_FVCORE = range(5)
SDB_BOX_LTRB, SDB_BOX_XYXY = 0, 1

# every symbol include/slender_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "sdb_last_error", "sdb_abi_version", "sdb_dcn_output_size", "sdb_dcn_supported",
    "sdb_dcn_workspace_bytes", "sdb_dcn_packed_input_bytes", "sdb_dcn_columns_bytes", "sdb_dcn_forward",
    "sdb_dcn_backward_data", "sdb_dcn_backward_weight", "sdb_dcn_prepared_weight_bytes", "sdb_dcn_prepare_weights",
    "sdb_dcn_multi_workspace_bytes", "sdb_dcn_forward_multi", "sdb_dcn_backward_multi", "sdb_assign_workspace_bytes",
    "sdb_iou_assign", "sdb_match_quality_assign", "sdb_pairwise_iou", "sdb_sigmoid_focal_loss",
    "sdb_box_reg_loss", "sdb_centerness_targets", "sdb_slender_centerness_targets", "sdb_fcos_location_targets_batched", "sdb_points_postprocess_workspace_bytes", "sdb_points_postprocess", "sdb_point_targets_workspace_bytes", "sdb_point_targets", "sdb_fcos_location_targets", "sdb_fcos_topk_workspace_bytes", "sdb_fcos_topk_location_targets", "sdb_reppoints_dcn_offset", "sdb_reppoints_dcn_offset_backward", "sdb_profile_enable", "sdb_profile_reset", "sdb_profile_read", "sdb_launch_count", "sdb_set_sm_reserve", "sdb_set_forward_pair", "sdb_set_backward_pair", "sdb_gn_relu_workspace_bytes", "sdb_gn_relu_forward", "sdb_gn_relu_backward",
]


class Geom(ctypes.Structure):
    """sdb_dcn_geom"""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "N", "C_in", "H", "W", "C_out", "kH", "kW", "sH", "sW", "pH", "pW", "dH", "dW", "groups",
        "deformable_groups")]


class Problem(ctypes.Structure):
    """sdb_dcn_problem: one row of a whole-head call (a FPN level of one convolution)"""
    _fields_ = [("N", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("weight_id", ctypes.c_int32),
                ("offset_group", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("x", ctypes.c_void_p), ("offset", ctypes.c_void_p), ("mask", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("x_packed", ctypes.c_void_p), ("grad_out", ctypes.c_void_p), ("grad_x", ctypes.c_void_p),
                ("grad_offset", ctypes.c_void_p), ("grad_mask", ctypes.c_void_p), ("columns", ctypes.c_void_p)]


class Weights(ctypes.Structure):
    """sdb_dcn_weights: one weight tensor of a whole-head call"""
    _fields_ = [("weight", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("prepared", ctypes.c_void_p),
                ("grad_weight", ctypes.c_void_p), ("grad_bias", ctypes.c_void_p)]


class GnTensor(ctypes.Structure):
    """sdb_gn_tensor"""
    _fields_ = [("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("grad_y", ctypes.c_void_p), ("grad_x", ctypes.c_void_p),
                ("stats", ctypes.c_void_p), ("N", ctypes.c_int32), ("HW", ctypes.c_int32), ("param_id", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class GnParams(ctypes.Structure):
    """sdb_gn_params"""
    _fields_ = [("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("grad_gamma", ctypes.c_void_p),
                ("grad_beta", ctypes.c_void_p)]


class PPLevel(ctypes.Structure):
    """sdb_pp_level: one FPN level of the inference post-processing"""
    _fields_ = [("cls", ctypes.c_void_p), ("pts", ctypes.c_void_p), ("centers", ctypes.c_void_p), ("H", ctypes.c_int32),
                ("W", ctypes.c_int32), ("stride", ctypes.c_float)]


SDB_MAX_PROBLEMS, SDB_MAX_WEIGHTS = 16, 4
SDB_BWD_WEIGHT_ONLY, SDB_BWD_DATA_ONLY, SDB_BWD_GRAD_PACKED, SDB_BWD_NO_GATHER, SDB_BWD_GATHER_ONLY = 1, 2, 4, 8, 16
SDB_BWD_BUILD_INDEX, SDB_BWD_INDEX_READY = 32, 64

_lib = None
_vp, _i32, _i64, _f32, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
_gp = ctypes.POINTER(Geom)


def _declare(lib):
    lib.sdb_last_error.restype = ctypes.c_char_p
    lib.sdb_last_error.argtypes = []
    lib.sdb_abi_version.restype = ctypes.c_int
    lib.sdb_dcn_output_size.argtypes = [_gp, ctypes.POINTER(_i32), ctypes.POINTER(_i32)]
    lib.sdb_dcn_supported.argtypes = [_gp, ctypes.c_int, ctypes.c_int]
    lib.sdb_dcn_workspace_bytes.restype = _sz
    lib.sdb_dcn_workspace_bytes.argtypes = [ctypes.c_int, _gp, ctypes.c_int, ctypes.c_int]
    lib.sdb_dcn_packed_input_bytes.restype = _sz
    lib.sdb_dcn_packed_input_bytes.argtypes = [_gp, ctypes.c_int]
    lib.sdb_dcn_columns_bytes.restype = _sz
    lib.sdb_dcn_columns_bytes.argtypes = [_gp, ctypes.c_int]
    lib.sdb_dcn_forward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _gp, ctypes.c_int, ctypes.c_int, _vp, _sz, _vp, _vp]
    lib.sdb_dcn_backward_data.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _gp, ctypes.c_int,
                                          ctypes.c_int, _vp, _sz, _vp, _vp]
    lib.sdb_dcn_backward_weight.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _gp, ctypes.c_int,
                                            ctypes.c_int, _vp, _sz, _vp, _vp]
    pp, wp = ctypes.POINTER(Problem), ctypes.POINTER(Weights)
    lib.sdb_dcn_prepared_weight_bytes.restype = _sz
    lib.sdb_dcn_prepared_weight_bytes.argtypes = [_gp, ctypes.c_int, ctypes.c_int]
    lib.sdb_dcn_prepare_weights.argtypes = [_vp, _vp, _gp, ctypes.c_int, ctypes.c_int, _vp, _vp]
    lib.sdb_dcn_multi_workspace_bytes.restype = _sz
    lib.sdb_dcn_multi_workspace_bytes.argtypes = [pp, _i32, wp, _i32, _gp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.sdb_dcn_forward_multi.argtypes = [pp, _i32, wp, _i32, _gp, ctypes.c_int, ctypes.c_int, _vp, _sz, _vp]
    lib.sdb_dcn_backward_multi.argtypes = [pp, _i32, wp, _i32, _gp, ctypes.c_int, ctypes.c_int, _f32, ctypes.c_int, _vp, _sz, _vp]
    lib.sdb_points_postprocess_workspace_bytes.restype = _sz
    lib.sdb_points_postprocess_workspace_bytes.argtypes = [_i32, _i32, _i32, _f32]
    lib.sdb_points_postprocess.argtypes = [ctypes.POINTER(PPLevel), _i32, _i32, _i32, _i32, _i32, ctypes.POINTER(_f32),
                                           ctypes.POINTER(_i32), _f32, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.sdb_assign_workspace_bytes.restype = _sz
    lib.sdb_assign_workspace_bytes.argtypes = [_i32, _i32, _i32]
    lib.sdb_iou_assign.argtypes = [_vp, _vp, _i32, _i32, ctypes.POINTER(_f32), ctypes.POINTER(ctypes.c_int8),
                                   _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.sdb_match_quality_assign.argtypes = [_vp, _i32, _i32, ctypes.POINTER(_f32), ctypes.POINTER(ctypes.c_int8),
                                             _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]
    lib.sdb_pairwise_iou.argtypes = [_vp, _vp, _i32, _i32, _vp, _vp]
    lib.sdb_sigmoid_focal_loss.argtypes = [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _vp, _vp, _vp]
    lib.sdb_box_reg_loss.argtypes = [_vp, _vp, _vp, _i64, ctypes.c_int, ctypes.c_int, _f32, _f32, _vp, _vp, _vp]
    lib.sdb_centerness_targets.argtypes = [_vp, _i64, _vp, _vp]
    lib.sdb_slender_centerness_targets.argtypes = [_vp, _i64, _vp, _vp]
    lib.sdb_fcos_location_targets_batched.argtypes = [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, ctypes.POINTER(_i32),
                                                      ctypes.POINTER(_f32), _i32, _f32, _i64, _i32, _i32, _vp, _vp, _vp,
                                                      _vp, _sz, _vp]
    lib.sdb_fcos_location_targets.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, ctypes.POINTER(_i32), ctypes.POINTER(_f32), _i32, _f32, _i64, _vp, _vp, _vp]
    lib.sdb_fcos_topk_workspace_bytes.restype = _sz
    lib.sdb_fcos_topk_workspace_bytes.argtypes = [_i32]
    lib.sdb_fcos_topk_location_targets.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, ctypes.POINTER(_i32), ctypes.POINTER(_f32), _i32, _f32, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]
    lib.sdb_point_targets_workspace_bytes.restype = _sz
    lib.sdb_point_targets_workspace_bytes.argtypes = [_i32]
    lib.sdb_point_targets.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _i64, _vp, _vp, _vp, _sz, _vp]
    lib.sdb_reppoints_dcn_offset.argtypes = [_vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp]
    lib.sdb_reppoints_dcn_offset_backward.argtypes = [_vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp]
    lib.sdb_launch_count.restype = ctypes.c_longlong
    lib.sdb_launch_count.argtypes = []
    lib.sdb_gn_relu_workspace_bytes.restype = _sz
    lib.sdb_gn_relu_workspace_bytes.argtypes = [ctypes.POINTER(GnTensor), _i32, _i32, _i32]
    for f in (lib.sdb_gn_relu_forward, lib.sdb_gn_relu_backward):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.POINTER(GnTensor), _i32, ctypes.POINTER(GnParams), _i32, _i32, _i32, _f32, _i32, _i32, _vp, _sz, _vp]
    lib.sdb_set_backward_pair.restype = ctypes.c_int
    lib.sdb_set_backward_pair.argtypes = [_i32]
    lib.sdb_set_forward_pair.restype = ctypes.c_int
    lib.sdb_set_forward_pair.argtypes = [_i32]
    lib.sdb_set_sm_reserve.restype = ctypes.c_int
    lib.sdb_set_sm_reserve.argtypes = [_i32]
    lib.sdb_profile_enable.argtypes = [ctypes.c_int]
    lib.sdb_profile_read.argtypes = [ctypes.c_int, ctypes.POINTER(_f32), ctypes.POINTER(ctypes.c_int)]
    return lib


def lib():
    """Load (once) and return the C-ABI library; loud failure when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libslender_b200.so is not built (%s missing). Build it with "
                "`python -m slenderobjdet_b200.csrc.build` (needs nvcc, targets sm_100a). "
                "There is no CPU or PyTorch fallback for these ops." % LIB_PATH)
        _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("slender_b200: " + lib().sdb_last_error().decode("utf-8", "replace"))


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def addr(t):
    """device address of a tensor as an int for a ctypes struct field (None -> NULL)"""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def io_dtype(t):
    if t.dtype == torch.float32:
        return SDB_F32
    if t.dtype == torch.bfloat16:
        return SDB_BF16
    raise RuntimeError("slender_b200: unsupported tensor dtype %s (float32 / bfloat16 only)" % t.dtype)
