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
    assert _lib.lib().sdb_abi_version() == 2


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
