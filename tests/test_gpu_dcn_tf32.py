"""GPU parity of the kind::tf32 tensor-core forward (SDB_MATH_TF32 / SDB_MATH_TF32X3, csrc/dcn_tf32.cu) against the
CPU oracle (fp32 sampling in the reference's operation order, double-accumulated contraction).

Tolerances (relative L2, plus the element-wise and per-128-pixel-block bounds of test_gpu_dcn_large.assert_close):
  tf32   : 1e-3   one pass, operands rounded to 10 mantissa bits (measured ~3e-4; BASELINE.json's 1e-4 is not
                  reachable with single-pass tf32 operands, which is why tf32x3 exists)
  tf32x3 : 5e-5   three error-compensated passes (measured 1.5e-5 at K = 2304: the accumulator's fp32 additions, not the
                  operands), inside BASELINE.json's "<= 1e-4 for tf32/fp32"
The backward of both modes runs the exact fp32 kernels: gradients are held to the fp32 tolerance 1e-4.
Reference semantics: detectron2/detectron2/layers/csrc/deformable/deform_conv_cuda_kernel.cu:96-130, :216-288, :785-868.
"""
import numpy as np
import pytest
import torch

import slenderobjdet_b200 as sdb
from slenderobjdet_b200 import _lib
from test_gpu_dcn_large import assert_close, compare, make_case, oracle, run_gpu

pytestmark = pytest.mark.gpu
TOL = {"tf32": 1e-3, "tf32x3": 5e-5}

# (N, C, H, W, O, sigma, kernel, stride, pad, dil)
SHAPES = [
    (2, 32, 13, 21, 16, 2.0, 3, 1, 1, 1),        # the smallest channel counts the path takes
    (1, 96, 25, 42, 80, 0.5, 3, 1, 1, 1),        # C not a multiple of 64, one accumulator half of 80 columns
    (2, 64, 20, 19, 192, 8.0, 3, 1, 1, 1),       # two halves: 128 + 64 columns
    (2, 256, 25, 42, 256, 2.0, 3, 1, 1, 1),      # the head's geometry (P5)
    (2, 128, 31, 29, 256, 30.0, 3, 1, 1, 1),     # most taps outside the image
    (2, 64, 33, 27, 64, 2.0, 3, 2, 1, 1),        # stride 2
    (1, 64, 30, 30, 128, 2.0, 3, 1, 2, 2),       # dilation 2
    (2, 128, 16, 24, 128, 1.0, 1, 1, 0, 1),      # 1x1
    (1, 64, 18, 22, 96, 2.0, (2, 4), 1, 1, 1),   # non-square kernel, 8 taps
]


@pytest.mark.parametrize("math", ["tf32", "tf32x3"])
@pytest.mark.parametrize("modulated", [False, True], ids=["v1", "v2"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "N%dC%d_%dx%d_O%d" % s[:5])
def test_tf32_forward_vs_oracle(math, modulated, shape):
    N, C, H, W, O, sigma, k, stride, pad, dil = shape
    c = make_case(500 + C + O, N, C, H, W, O, modulated, sigma=sigma, k=k, stride=stride, pad=pad, dil=dil, wscale=0.05)
    got = run_gpu(c, torch.float32, backward=False, math=math)
    ref = oracle(c, backward=False)
    errs = compare(got, ref, TOL[math], "%s %s" % (math, shape,))
    print(math, shape, errs)


@pytest.mark.parametrize("math", ["tf32", "tf32x3"])
def test_tf32_benchmarked_p3_map(math):
    """2 x 256 x 100 x 168 -> 263 tiles on 148 persistent CTAs: second tile per CTA, accumulator flip, ring parities."""
    c = make_case(1168, 2, 256, 100, 168, 256, False, sigma=2.0, wscale=0.01)
    got = run_gpu(c, torch.float32, backward=False, math=math)
    ref = oracle(c, backward=False)
    errs = compare(got, ref, TOL[math], math + " P3")
    print(math, errs)


@pytest.mark.parametrize("math", ["tf32", "tf32x3"])
def test_tf32_modes_keep_fp32_gradients(math):
    """forward on tensor cores, backward on the exact fp32 kernels: all gradients at the fp32 tolerance"""
    c = make_case(77, 2, 64, 17, 23, 64, True, sigma=2.0, wscale=0.05)
    got = run_gpu(c, torch.float32, backward=True, math=math)
    ref = oracle(c, backward=True)
    assert_close(got["out"], ref["out"], TOL[math], "out")
    for k in ("grad_x", "grad_offset", "grad_mask", "grad_weight", "grad_bias"):
        assert_close(got[k], ref[k], 1e-4, k)


def test_tf32_zero_offset_equals_conv2d():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 19, 23, generator=g).cuda()
    w = (torch.randn(96, 64, 3, 3, generator=g) * 0.05).cuda()
    off = torch.zeros(2, 18, 19, 23, device="cuda")
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1).float()
    for math, tol in TOL.items():
        with sdb.dcn_math(math):
            y = sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
        assert_close(y.cpu().numpy(), ref.cpu().numpy(), tol, math)


def test_tf32_multi_call_and_prepared_weights():
    """whole-head call (5 levels x 2 convolutions, shared prepared images) == per-problem calls, bit for bit"""
    g = torch.Generator().manual_seed(9)
    levels = [(20, 34), (10, 17), (5, 9), (3, 5), (2, 3)]
    ws = [(torch.randn(64, 64, 3, 3, generator=g) * 0.05).cuda() for _ in range(2)]
    xs, offs, wids = [], [], []
    for (H, W) in levels:
        off = (torch.randn(2, 18, H, W, generator=g) * 2).cuda()
        for k in range(2):
            xs.append(torch.randn(2, 64, H, W, generator=g).cuda())
            offs.append(off)
            wids.append(k)
    with sdb.dcn_math("tf32x3"):
        outs = sdb.deform_conv_multi(xs, offs, ws, 1, 1, 1, weight_ids=wids)
        for i, o in enumerate(outs):
            single = sdb.deform_conv(xs[i], offs[i], ws[wids[i]], 1, 1, 1, 1, 1)
            assert torch.equal(o, single), i


def test_tf32_unsupported_geometry_raises():
    x = torch.randn(1, 24, 8, 8, device="cuda")
    w = torch.randn(16, 24, 3, 3, device="cuda")
    off = torch.zeros(1, 18, 8, 8, device="cuda")
    with sdb.dcn_math("tf32"), pytest.raises(RuntimeError, match="C_in not a multiple of 32"):
        sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
    gm = _lib.Geom(1, 64, 8, 8, 64, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    import ctypes
    assert _lib.lib().sdb_dcn_supported(ctypes.byref(gm), _lib.SDB_BF16, _lib.SDB_MATH_TF32) == 0   # float32 tensors only
    assert _lib.lib().sdb_dcn_supported(ctypes.byref(gm), _lib.SDB_F32, _lib.SDB_MATH_TF32X3) == 1
