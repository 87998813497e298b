"""CPU: the C-ABI library loads and exports every symbol include/slender_b200.h declares; host-side
API mirrors the reference's error behaviour.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
import slenderobjdet_b200 as sdb
from slenderobjdet_b200 import _lib
from slenderobjdet_b200.layers.deform_conv import _DeformConv


def _header_functions():
    src = open(os.path.join(ROOT, "include", "slender_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run python -m slenderobjdet_b200.csrc.build"
    assert _lib.lib().sdb_abi_version() == 4


def test_exports_every_declared_symbol():
    declared = _header_functions()
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sdb_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert getattr(handle, s) is not None


def test_no_libcuda_or_torch_link_dependency():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libc10" not in out


def test_geometry_validation_without_gpu():
    lib = _lib.lib()
    g = _lib.Geom(2, 256, 100, 152, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    assert lib.sdb_dcn_output_size(ctypes.byref(g), ho, wo) == 0 and (ho.value, wo.value) == (100, 152)
    bad = _lib.Geom(1, 4, 2, 2, 4, 5, 5, 1, 1, 0, 0, 1, 1, 1, 1)  # output would be <= 0
    assert lib.sdb_dcn_output_size(ctypes.byref(bad), ho, wo) == -1
    assert b"too small" in lib.sdb_last_error()
    bad = _lib.Geom(1, 6, 8, 8, 4, 3, 3, 1, 1, 1, 1, 1, 1, 4, 1)  # channels % groups
    assert lib.sdb_dcn_output_size(ctypes.byref(bad), ho, wo) == -1
    assert lib.sdb_dcn_workspace_bytes(0, ctypes.byref(g), _lib.SDB_F32, _lib.SDB_MATH_FP32) == 0


def test_python_errors_match_reference():
    m = sdb.DeformConv(4, 4, 3, padding=1)
    with pytest.raises(NotImplementedError):  # deform_conv.py:48-49
        m(torch.randn(1, 4, 5, 5), torch.zeros(1, 18, 5, 5))
    with pytest.raises(ValueError):  # deform_conv.py:29-32
        sdb.deform_conv(torch.randn(4, 5, 5), torch.zeros(1, 18, 5, 5), m.weight)
    with pytest.raises(ValueError):  # deform_conv.py:147-152
        sdb.deform_conv(torch.randn(1, 4, 1, 1), torch.zeros(1, 18, 1, 1), torch.randn(4, 4, 5, 5))
    with pytest.raises(AssertionError):  # deform_conv.py:335
        sdb.DeformConv(4, 4, 3, bias=True)
    m2 = sdb.ModulatedDeformConv(4, 4, 3, padding=1)
    with pytest.raises(NotImplementedError):
        m2(torch.randn(1, 4, 5, 5), torch.zeros(1, 18, 5, 5), torch.ones(1, 9, 5, 5))


def test_empty_input_shortcut():
    m = sdb.DeformConv(4, 6, 3, padding=1)
    y = m(torch.zeros(0, 4, 7, 9), torch.zeros(0, 18, 7, 9))  # deform_conv.py:362-374
    assert tuple(y.shape) == (0, 6, 7, 9)


def test_im2col_step_rule():
    f = _DeformConv._cal_im2col_step  # deform_conv.py:155-176
    assert f(2, 64) == 2 and f(8, 64) == 8 and f(16, 64) == 16 and f(128, 64) == 64 and f(130, 64) == 26
    assert f(97, 64) == 1


def test_state_dict_contract():
    assert list(sdb.DeformConv(8, 6, 3).state_dict().keys()) == ["weight"]
    assert tuple(sdb.DeformConv(8, 6, 3, groups=2).weight.shape) == (6, 4, 3, 3)
    assert list(sdb.ModulatedDeformConv(8, 6, 3).state_dict().keys()) == ["weight", "bias"]
    assert list(sdb.DFConv2d(8, 8).state_dict().keys()) == ["offset.weight", "offset.bias", "conv.weight"]
    assert sdb.DFConv2d(8, 8).offset.out_channels == 27 and sdb.DFConv2d(8, 8, with_modulated_dcn=False).offset.out_channels == 18
    r = repr(sdb.ModulatedDeformConv(8, 6, 3, padding=1))
    assert "deformable_groups=1, bias=True" in r


def test_matcher_constructor_contract():
    t = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=10)
    assert t.thresholds == [-float("inf"), 0.3, 0.7, float("inf")] and t.topk == 10
    with pytest.raises(AssertionError):
        sdb.Matcher([0.0, 0.7], [0, -1, 1])
    with pytest.raises(RuntimeError):
        t(torch.zeros(3, 20))  # CPU tensors: no fallback


def _table(levels, batch=2, groups=True):
    rows = []
    for li, (h, w) in enumerate(levels):
        for b in range(2):
            rows.append(_lib.Problem(batch, h, w, b, li if groups else -1, 0, 0x1000, 0x2000 + li, None, 0x3000, None,
                                     0x4000, 0x5000, 0x6000, None))
    return (_lib.Problem * len(rows))(*rows), len(rows)


def test_multi_problem_table_host_logic_without_gpu():
    """sdb_dcn_multi_workspace_bytes plans the workspace on the host: sizes, sharing of the transposed index between
    the two convolutions of a level, and table validation -- no kernel is launched."""
    lib = _lib.lib()
    g = _lib.Geom(1, 256, 8, 8, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)    # N / H / W of the common geometry are ignored
    gp = ctypes.byref(g)
    wts = (_lib.Weights * 2)(_lib.Weights(0x7000, None, None, 0x8000, None), _lib.Weights(0x7100, None, None, 0x8100, None))
    levels = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    probs, n = _table(levels)
    fwd = lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 0)
    bwd = lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 1)
    px = 2 * sum(h * w for h, w in levels)
    assert fwd >= 2 * px * 256 * 2          # an NHWC bf16 copy of every input
    assert bwd > fwd
    # both branches of a level share one transposed index: without the groups the plan needs one per problem
    probs_ng, _ = _table(levels, groups=False)
    bwd_ng = lib.sdb_dcn_multi_workspace_bytes(probs_ng, n, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 1)
    assert bwd_ng > bwd + px * 9 * 32 // 2
    # prepared weights passed by the caller are not planned into the workspace
    wts_p = (_lib.Weights * 2)(_lib.Weights(0x7000, None, 0x9000, 0x8000, None), _lib.Weights(0x7100, None, 0x9100, 0x8100, None))
    assert lib.sdb_dcn_multi_workspace_bytes(probs, n, wts_p, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 0) == \
        fwd - 2 * lib.sdb_dcn_prepared_weight_bytes(gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16)
    # fp32 math needs no workspace; a bad weight_id, too many problems, or a group whose members differ are refused
    assert lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, 2, gp, _lib.SDB_F32, _lib.SDB_MATH_FP32, 1) == 0
    probs[3].weight_id = 5
    assert lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 1) == 0
    assert b"weight_id" in lib.sdb_last_error()
    probs[3].weight_id = 1
    probs[1].offset = 0x2fff                 # same group as problem 0, different offset tensor
    assert lib.sdb_dcn_multi_workspace_bytes(probs, n, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 1) == 0
    assert b"offset_group" in lib.sdb_last_error()
    big, nb = _table(levels * 2)
    assert nb == 20 and lib.sdb_dcn_multi_workspace_bytes(big, nb, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 0) == 0
    # the single-problem entry points are one-row tables
    g1 = _lib.Geom(2, 256, 100, 168, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    one = lib.sdb_dcn_workspace_bytes(_lib.SDB_OP_BACKWARD_DATA, ctypes.byref(g1), _lib.SDB_BF16, _lib.SDB_MATH_BF16)
    assert one > 0 and one == lib.sdb_dcn_workspace_bytes(_lib.SDB_OP_BACKWARD_WEIGHT, ctypes.byref(g1), _lib.SDB_BF16, _lib.SDB_MATH_BF16)
    assert one > 2 * 100 * 168 * 9 * 256 * 2      # a DEFORMABLE convolution's plan: holds the dcol tiles (taps * C bf16 per pixel)
    # offset == NULL in every problem = plain convolution (tower Conv2d): no dcol tiles, no transposed index; mixed tables
    # are refused
    conv, nc = _table(levels)
    for i in range(nc):
        conv[i].offset = None
        conv[i].grad_offset = None
        conv[i].offset_group = -1
    bwd_conv = lib.sdb_dcn_multi_workspace_bytes(conv, nc, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 1)
    assert 0 < bwd_conv < bwd - px * 9 * 256 * 2 // 2
    conv[2].offset = 0x2000
    assert lib.sdb_dcn_multi_workspace_bytes(conv, nc, wts, 2, gp, _lib.SDB_BF16, _lib.SDB_MATH_BF16, 1) == 0
    assert b"plain convolution" in lib.sdb_last_error()


def test_deform_conv_multi_argument_checks():
    w = torch.zeros(4, 4, 3, 3)
    with pytest.raises(ValueError):
        sdb.deform_conv_multi([torch.zeros(1, 4, 5, 5)], [], w)
    with pytest.raises(ValueError):
        sdb.deform_conv_multi([torch.zeros(1, 4, 5, 5)] * 17, [torch.zeros(1, 18, 5, 5)] * 17, w)
    with pytest.raises(NotImplementedError):   # CPU tensors, like the reference (deform_conv.py:48-49)
        sdb.deform_conv_multi([torch.zeros(1, 4, 5, 5)], [torch.zeros(1, 18, 5, 5)], w, 1, 1, 1)
    assert sdb.deform_conv_multi([], [], w) == []


def test_backward_flag_validation_without_gpu():
    """sdb_dcn_backward_multi refuses contradictory phase flags before touching the device (flags are validated first)."""
    lib = _lib.lib()
    g = _lib.Geom(1, 256, 8, 8, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    probs, n = _table([(25, 42)])
    wts = (_lib.Weights * 2)(_lib.Weights(0x7000, None, None, 0x8000, None), _lib.Weights(0x7100, None, None, 0x8100, None))
    L = _lib
    bad = [L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_DATA_ONLY, L.SDB_BWD_NO_GATHER | L.SDB_BWD_GATHER_ONLY,
           L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_GATHER_ONLY, L.SDB_BWD_BUILD_INDEX, L.SDB_BWD_BUILD_INDEX | L.SDB_BWD_DATA_ONLY,
           L.SDB_BWD_INDEX_READY | L.SDB_BWD_WEIGHT_ONLY, L.SDB_BWD_INDEX_READY | L.SDB_BWD_NO_GATHER, 128]
    for f in bad:
        rc = lib.sdb_dcn_backward_multi(probs, n, wts, 2, ctypes.byref(g), L.SDB_BF16, L.SDB_MATH_BF16, ctypes.c_float(1.0), f,
                                        None, 0, None)
        assert rc != 0, f
        msg = lib.sdb_last_error()
        assert b"bad backward flags" in msg, (f, msg)
