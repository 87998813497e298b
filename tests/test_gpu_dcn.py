"""GPU parity: CUDA deformable convolution (through the Python operator API -> ctypes -> C ABI)
against the CPU oracle / committed goldens.  Tolerances from BASELINE.json north_star:
rel <= 1e-4 for the fp32 path, rel <= 1e-2 for bf16 tensor-core math."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
import slenderobjdet_b200 as sdb
from oracle import dcn as odcn

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _dev(a, dtype=torch.float32):
    return None if a is None else torch.as_tensor(a).to("cuda", dtype)


def _cfg(c):
    sh, sw, ph, pw, dh, dw, g, dg = [int(v) for v in c["cfg"]]
    return (sh, sw), (ph, pw), (dh, dw), g, dg


def _run(c, math, need_grads=True):
    st, pd, dl, g, dg = _cfg(c)
    x, off, w = _dev(c["x"]).requires_grad_(), _dev(c["offset"]).requires_grad_(), _dev(c["weight"]).requires_grad_()
    m = _dev(c.get("mask"))
    b = _dev(c.get("bias"))
    with sdb.dcn_math(math):
        if m is None:
            y = sdb.deform_conv(x, off, w, st, pd, dl, g, dg)
        else:
            assert st[0] == st[1] and pd[0] == pd[1] and dl[0] == dl[1]
            m.requires_grad_()
            if b is not None:
                b.requires_grad_()
            y = sdb.modulated_deform_conv(x, off, m, w, b, st[0], pd[0], dl[0], g, dg)
        if need_grads:
            y.backward(_dev(c["grad_out"]))
    torch.cuda.synchronize()
    return y, x, off, w, m, b


def test_known_answer_reference_test():
    """/root/reference/tests/test_deformable_conv.py:85-87 on the CUDA kernels (fp32 math)."""
    z = np.load(f"{GOLDEN}/dcn_known_answer.npz")
    conv = sdb.DeformConv(2, 1, 3, 1, 1, bias=False).cuda()
    conv.weight.data = _dev(z["weight"])
    with sdb.dcn_math("fp32"):
        y1 = conv(_dev(z["x"]), _dev(z["offsets_1"])).detach().cpu().numpy()
        y2 = conv(_dev(z["x"]), _dev(z["offsets_2"])).detach().cpu().numpy()
    assert np.all(np.abs(y2 - z["expected_conv"]) < 1e-5)
    assert np.all(np.abs(y2 - z["expected_dconv_zero"]) < 1e-5)
    assert np.all(np.abs(y1 - z["expected_dconv_grid"]) < 1e-5)


CASES = ["v1_basic", "v1_big_offsets", "v1_groups_dg", "v1_stride2_dil2", "v1_k1", "v1_k5x3", "v1_c64",
         "v2_basic", "v2_nobias_dg2", "v2_c64"]


@pytest.mark.parametrize("name", CASES)
def test_fp32_path_vs_golden(dcn_cases, name):
    c = dcn_cases[name]
    y, x, off, w, m, b = _run(c, "fp32")
    tol = TOL["fp32"]
    assert rel_err(y.detach().cpu().numpy(), c["out"]) < tol
    assert rel_err(x.grad.cpu().numpy(), c["grad_x"]) < tol
    assert rel_err(off.grad.cpu().numpy(), c["grad_offset"]) < tol
    assert rel_err(w.grad.cpu().numpy(), c["grad_weight"]) < tol
    if m is not None:
        assert rel_err(m.grad.cpu().numpy(), c["grad_mask"]) < tol
    if b is not None:
        assert rel_err(b.grad.cpu().numpy(), c["grad_bias"]) < tol


def _random_case(seed, N, C, H, W, O, modulated, sigma, k=3, pad=1):
    g = torch.Generator().manual_seed(seed)
    c = dict(x=torch.randn(N, C, H, W, generator=g).numpy(),
             weight=(torch.randn(O, C, k, k, generator=g) * 0.05).numpy(),
             offset=(torch.randn(N, 2 * k * k, H, W, generator=g) * sigma).numpy(),
             grad_out=torch.randn(N, O, H, W, generator=g).numpy(),
             cfg=np.array([1, 1, pad, pad, 1, 1, 1, 1]))
    if modulated:
        c["mask"] = torch.sigmoid(torch.randn(N, k * k, H, W, generator=g)).numpy()
        c["bias"] = torch.randn(O, generator=g).numpy()
    return c


def _oracle(c):
    st, pd, dl, g, dg = _cfg(c)
    kw = dict(stride=st, padding=pd, dilation=dl, groups=g, deformable_groups=dg)
    y = odcn.forward(c["x"], c["offset"], c["weight"], mask=c.get("mask"), bias=c.get("bias"), **kw)
    gr = odcn.backward(c["x"], c["offset"], c["weight"], c["grad_out"], mask=c.get("mask"),
                       with_bias="bias" in c, **kw)
    return y, gr


@pytest.mark.parametrize("math", ["fp32", "bf16"])
@pytest.mark.parametrize("modulated", [False, True])
@pytest.mark.parametrize("shape", [(2, 64, 13, 21, 64, 2.0), (1, 128, 25, 42, 256, 0.5), (2, 256, 7, 11, 256, 8.0),
                                   (3, 256, 20, 19, 128, 30.0)])
def test_vs_oracle_random(math, modulated, shape):
    N, C, H, W, O, sigma = shape
    c = _random_case(77 + N + C, N, C, H, W, O, modulated, sigma)
    y, x, off, w, m, b = _run(c, math)
    yo, go = _oracle(c)
    tol = TOL[math]
    assert rel_err(y.detach().cpu().numpy(), yo) < tol
    assert rel_err(x.grad.cpu().numpy(), go["grad_x"]) < tol
    assert rel_err(off.grad.cpu().numpy(), go["grad_offset"]) < tol
    assert rel_err(w.grad.cpu().numpy(), go["grad_weight"]) < tol
    if modulated:
        assert rel_err(m.grad.cpu().numpy(), go["grad_mask"]) < tol
        assert rel_err(b.grad.cpu().numpy(), go["grad_bias"]) < tol


@pytest.mark.parametrize("math", ["fp32", "bf16"])
def test_zero_offset_equals_conv2d(math):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 17, 23, generator=g).cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).cuda()
    with sdb.dcn_math(math):
        y = sdb.deform_conv(x, torch.zeros(2, 18, 17, 23, device="cuda"), w, 1, 1, 1, 1, 1)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL[math]


@pytest.mark.parametrize("shape", [(2, 64, 12, 14, 64), (2, 128, 13, 21, 80), (1, 64, 25, 42, 256), (3, 64, 7, 11, 48),
                                   (2, 256, 50, 84, 256), (1, 64, 20, 24, 72), (2, 64, 9, 36, 64)])
def test_bf16_tensors_io(shape):
    """bf16 tensors in and out (the layout-pack kernels have bf16-specialised paths: even / odd plane sizes,
    ragged 8x16 patches at the borders, C_out not a multiple of 64)."""
    N, C, H, W, O = shape
    c = _random_case(9 + H, N, C, H, W, O, False, 1.0)
    # the oracle gets the bf16-rounded tensors the kernel is handed, so the 1e-2 bar measures the kernel alone
    q = lambda a: torch.as_tensor(a).bfloat16().float().numpy()
    cq = dict(c, x=q(c["x"]), weight=q(c["weight"]), grad_out=q(c["grad_out"]))
    yo, go = _oracle(cq)
    x = _dev(c["x"], torch.bfloat16).requires_grad_()
    off = _dev(c["offset"]).requires_grad_()
    w = _dev(c["weight"], torch.bfloat16).requires_grad_()
    y = sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
    assert y.dtype == torch.bfloat16
    y.backward(_dev(c["grad_out"], torch.bfloat16))
    assert x.grad.dtype == torch.bfloat16 and w.grad.dtype == torch.bfloat16 and off.grad.dtype == torch.float32
    assert rel_err(y.float().detach().cpu().numpy(), yo) < TOL["bf16"]
    assert rel_err(x.grad.float().cpu().numpy(), go["grad_x"]) < TOL["bf16"]
    assert rel_err(w.grad.float().cpu().numpy(), go["grad_weight"]) < TOL["bf16"]
    assert rel_err(off.grad.cpu().numpy(), go["grad_offset"]) < TOL["bf16"]


def test_all_taps_outside():
    x = torch.randn(1, 64, 9, 9, device="cuda", requires_grad=True)
    w = torch.randn(64, 64, 3, 3, device="cuda", requires_grad=True)
    off = torch.full((1, 18, 9, 9), 1000.0, device="cuda", requires_grad=True)
    for math in ("fp32", "bf16"):
        with sdb.dcn_math(math):
            y = sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
            y.sum().backward()
        assert float(y.abs().max()) == 0.0
        assert float(x.grad.abs().max()) == 0.0 and float(off.grad.abs().max()) == 0.0
        x.grad = None; off.grad = None; w.grad = None


@pytest.mark.parametrize("spread", [0.0, 0.6, 3.0])
def test_bf16_grad_input_with_long_transposed_lists(spread):
    """grad_input is computed as a gather over the transposed sampling index; offsets that send every
    tap of every pixel to (nearly) the same input location give lists of hundreds of entries per
    (input pixel, tap) - far beyond the 4-entry descriptor and the staged overflow descriptors - while
    most input pixels get empty lists."""
    N, C, H, W, O = 2, 64, 11, 13, 64
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(O, C, 3, 3, generator=g) * 0.05
    gy = torch.randn(N, O, H, W, generator=g)
    off = torch.zeros(N, 18, H, W)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    for i in range(3):
        for j in range(3):
            t = i * 3 + j
            off[:, 2 * t] = (H / 2 + 0.3) - (ys - 1 + i)       # every tap lands at (H/2+0.3, W/2+0.6) ...
            off[:, 2 * t + 1] = (W / 2 + 0.6) - (xs - 1 + j)
    off += torch.randn(N, 18, H, W, generator=g) * spread        # ... plus a spread
    c = dict(x=x.numpy(), offset=off.numpy(), weight=w.numpy(), grad_out=gy.numpy(), cfg=np.array([1, 1, 1, 1, 1, 1, 1, 1]))
    y, xd, offd, wd, _, _ = _run(c, "bf16")
    yo, go = _oracle(c)
    assert rel_err(y.detach().cpu().numpy(), yo) < TOL["bf16"]
    # grad_input sums each transposed list into a bf16 GEMM operand in an order set by atomics; the tail of very long
    # lists is accumulated in fp32 (dcn_tc.cu), which keeps the ~140-entry case at 4.4e-3 .. 5.9e-3 from run to run
    # (scratch/flaky_dx.py, 40 runs; 7.3e-3 .. 9.6e-3 with a bf16 read-modify-write per descriptor)
    assert rel_err(xd.grad.cpu().numpy(), go["grad_x"]) < TOL["bf16"]
    assert rel_err(offd.grad.cpu().numpy(), go["grad_offset"]) < TOL["bf16"]
    assert rel_err(wd.grad.cpu().numpy(), go["grad_weight"]) < TOL["bf16"]


def test_shape_errors_are_runtime_errors():
    x = torch.randn(1, 8, 6, 6, device="cuda")
    w = torch.randn(4, 8, 3, 3, device="cuda")
    with pytest.raises(RuntimeError):  # wrong offset channels (deform_conv_cuda.cu:242-244)
        sdb.deform_conv(x, torch.zeros(1, 16, 6, 6, device="cuda"), w, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):  # wrong offset spatial size (:236-240)
        sdb.deform_conv(x, torch.zeros(1, 18, 5, 6, device="cuda"), w, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):  # weight planes (:206-210)
        sdb.deform_conv(x, torch.zeros(1, 18, 6, 6, device="cuda"), torch.randn(4, 6, 3, 3, device="cuda"), 1, 1, 1, 1, 1)


def test_dfconv2d_and_grad_flow():
    torch.manual_seed(0)
    for mod in (False, True):
        m = sdb.DFConv2d(64, 64, with_modulated_dcn=mod, bias=mod).cuda()
        x = torch.randn(2, 64, 10, 12, device="cuda", requires_grad=True)
        y = m(x)
        y.square().mean().backward()
        assert tuple(y.shape) == (2, 64, 10, 12)
        for p in m.parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all()
        assert m.offset.weight.grad.abs().sum() > 0


@pytest.mark.parametrize("shape", [
    # N, C, H, W, O, k, stride, pad, dil, modulated, sigma
    (2, 64, 13, 21, 64, 3, 1, 1, 1, False, 2.0),
    (1, 128, 25, 42, 256, 3, 1, 1, 1, True, 0.5),
    (2, 256, 50, 84, 256, 3, 1, 1, 1, False, 2.0),
    (1, 256, 7, 11, 256, 3, 1, 1, 1, True, 8.0),
    (2, 192, 9, 10, 48, 3, 2, 1, 1, False, 1.0),
    (1, 64, 12, 9, 16, 1, 1, 0, 1, False, 1.0),
    (1, 512, 10, 13, 128, 3, 1, 2, 2, True, 3.0),
    (3, 256, 20, 19, 240, 3, 1, 1, 1, False, 30.0),
    (1, 64, 11, 12, 32, (2, 4), 1, 1, 1, False, 1.0),
])
def test_tensor_core_forward_vs_oracle(shape):
    """tcgen05 forward (bf16 operands, fp32 accumulate) vs the CPU oracle, rel <= 1e-2."""
    N, C, H, W, O, k, st, pd, dl, mod, sigma = shape
    kh, kw = (k, k) if isinstance(k, int) else k
    g = torch.Generator().manual_seed(C + O + H)
    Ho = (H + 2 * pd - (dl * (kh - 1) + 1)) // st + 1
    Wo = (W + 2 * pd - (dl * (kw - 1) + 1)) // st + 1
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(O, C, kh, kw, generator=g) * 0.05
    off = torch.randn(N, 2 * kh * kw, Ho, Wo, generator=g) * sigma
    m = torch.sigmoid(torch.randn(N, kh * kw, Ho, Wo, generator=g)) if mod else None
    b = torch.randn(O, generator=g) if mod else None
    yo = odcn.forward(x.numpy(), off.numpy(), w.numpy(), mask=None if m is None else m.numpy(),
                      bias=None if b is None else b.numpy(), stride=st, padding=pd, dilation=dl)
    with sdb.dcn_math("bf16"), torch.no_grad():
        if mod:
            y = sdb.modulated_deform_conv(x.cuda(), off.cuda(), m.cuda(), w.cuda(), b.cuda(), st, pd, dl, 1, 1)
        else:
            y = sdb.deform_conv(x.cuda(), off.cuda(), w.cuda(), st, pd, dl, 1, 1)
    torch.cuda.synchronize()
    assert rel_err(y.cpu().numpy(), yo) < TOL["bf16"]
    # against an oracle fed the SAME bf16-rounded operands the kernel uses, the gap is fp32-accumulation small
    yq = odcn.forward(x.bfloat16().float().numpy(), off.numpy(), w.bfloat16().float().numpy(),
                      mask=None if m is None else m.numpy(), bias=None if b is None else b.numpy(),
                      stride=st, padding=pd, dilation=dl)
    assert rel_err(y.cpu().numpy(), yq) < 6e-3  # bf16 interpolation + bf16 operand rounding of the sample


@pytest.mark.parametrize("save_columns", [True, False])
@pytest.mark.parametrize("shape", [
    # N, C, H, W, O, k, stride, pad, dil, modulated, sigma
    (2, 192, 9, 10, 48, 3, 2, 1, 1, False, 1.0),      # stride 2, three 64-channel chunks, C_out not a multiple of 64
    (1, 64, 12, 9, 16, 1, 1, 0, 1, False, 1.0),       # 1x1 kernel, no padding
    (1, 512, 10, 13, 128, 3, 1, 2, 2, True, 3.0),     # dilation 2, pad 2, four 128-channel chunks, mask + bias
    (3, 256, 20, 19, 240, 3, 1, 1, 1, False, 30.0),   # huge offsets (most taps outside), odd width
    (1, 64, 11, 12, 32, (2, 4), 1, 1, 1, False, 1.0),  # non-square kernel
    (2, 128, 16, 24, 80, 3, 2, 2, 2, True, 2.0),      # stride 2 + dilation 2, mask + bias
    (2, 256, 13, 21, 256, 3, 1, 0, 1, False, 2.0),    # no padding: output grid smaller than the input grid
])
def test_tensor_core_backward_geometries(shape, save_columns):
    """Every gradient of the tcgen05 backward (dcol GEMM + channel reduction, grad_input gather over the transposed
    index, weight gradient over saved columns or by re-sampling) away from 3x3 / stride 1 / pad 1, vs the CPU oracle."""
    N, C, H, W, O, k, st, pd, dl, mod, sigma = shape
    kh, kw = (k, k) if isinstance(k, int) else k
    g = torch.Generator().manual_seed(7 * C + O + H)
    Ho = (H + 2 * pd - (dl * (kh - 1) + 1)) // st + 1
    Wo = (W + 2 * pd - (dl * (kw - 1) + 1)) // st + 1
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(O, C, kh, kw, generator=g) * 0.05
    off = torch.randn(N, 2 * kh * kw, Ho, Wo, generator=g) * sigma
    m = torch.sigmoid(torch.randn(N, kh * kw, Ho, Wo, generator=g)) if mod else None
    b = torch.randn(O, generator=g) if mod else None
    gy = torch.randn(N, O, Ho, Wo, generator=g)
    ref = odcn.backward(x.numpy(), off.numpy(), w.numpy(), gy.numpy(), mask=None if m is None else m.numpy(),
                        with_bias=mod, stride=st, padding=pd, dilation=dl)
    xd, od, wd = x.cuda().requires_grad_(), off.cuda().requires_grad_(), w.cuda().requires_grad_()
    md = m.cuda().requires_grad_() if mod else None
    bd = b.cuda().requires_grad_() if mod else None
    sdb.set_dcn_save_columns(save_columns)
    try:
        with sdb.dcn_math("bf16"):
            if mod:
                y = sdb.modulated_deform_conv(xd, od, md, wd, bd, st, pd, dl, 1, 1)
            else:
                y = sdb.deform_conv(xd, od, wd, st, pd, dl, 1, 1)
            y.backward(gy.cuda())
        torch.cuda.synchronize()
    finally:
        sdb.set_dcn_save_columns(True)
    got = dict(grad_x=xd.grad, grad_offset=od.grad, grad_weight=wd.grad)
    if mod:
        got.update(grad_mask=md.grad, grad_bias=bd.grad)
    for name, t in got.items():
        e = rel_err(t.float().cpu().numpy(), ref[name])
        assert e < TOL["bf16"], (name, e)


def test_auto_mode_keeps_float32_tensors_exact():
    """'auto' follows the tensors: a float32 model gets the reference's fp32 numerics (rel <= 1e-4) even where the
    tensor-core geometry would qualify; bf16 autocast or bfloat16 tensors select the tcgen05 kernels."""
    c = _random_case(41, 2, 256, 13, 21, 256, False, 2.0)
    yo, _ = _oracle(c)
    x, off, w = _dev(c["x"]), _dev(c["offset"]), _dev(c["weight"])
    with sdb.dcn_math("auto"), torch.no_grad():
        y = sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
        assert rel_err(y.cpu().numpy(), yo) < TOL["fp32"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ya = sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
        e = rel_err(ya.float().cpu().numpy(), yo)
        assert 1e-4 < e < TOL["bf16"]          # tensor-core numerics: bf16 operands
    with pytest.raises(RuntimeError):
        sdb.deform_conv(x.double(), off, w.double(), 1, 1, 1, 1, 1)


def test_auto_mode_falls_back_to_fp32_kernels_for_unsupported_geometry():
    """5x5 taps / C_in=8 are outside the tensor-core path: 'auto' must run the exact fp32 kernels
    (still CUDA, never CPU) and 'bf16' must refuse loudly."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 9, 10, generator=g)
    w = torch.randn(16, 64, 5, 5, generator=g) * 0.05
    off = torch.randn(1, 50, 9, 10, generator=g)
    yo = odcn.forward(x.numpy(), off.numpy(), w.numpy(), stride=1, padding=2)
    with sdb.dcn_math("auto"), torch.no_grad():
        y = sdb.deform_conv(x.cuda(), off.cuda(), w.cuda(), 1, 2, 1, 1, 1)
    assert rel_err(y.cpu().numpy(), yo) < TOL["fp32"]
    with sdb.dcn_math("bf16"), pytest.raises(RuntimeError):
        sdb.deform_conv(x.cuda(), off.cuda(), w.cuda(), 1, 2, 1, 1, 1)
