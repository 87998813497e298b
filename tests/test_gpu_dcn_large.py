"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs[0], [1], [2], [4]) and for the tensor-core
backward outside the 3x3 / stride-1 envelope.

The persistent tcgen05 kernels launch min(tiles, 148) CTAs, so only maps with more than 148 tiles of 128 pixels
make a CTA walk a second tile (TMEM double-buffer flip, ring parities carried across tiles, weight re-stream,
overflow staging across tiles).  Every test here has 238 .. 2 100 tiles.  Besides the global relative L2 error
BASELINE.json states its tolerances in, every comparison carries
  * an element-wise bound  max|a-b| <= tol * max|b|   (a wrong border column or pixel cannot hide), and
  * a per-tile bound: relative L2 error of every block of 128 consecutive output positions <= 3 * tol
    (one bad tile in 263 moves the global norm by 0.4 % -- it cannot hide here).
Reference semantics: detectron2/detectron2/layers/csrc/deformable/deform_conv_cuda_kernel.cu:216-452, :785-1066;
shapes of /root/reference/tests/test_deformable_conv.py:69-87 scaled to the FPN maps of SURVEY.md section 8.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
import slenderobjdet_b200 as sdb
from oracle import dcn as odcn

pytestmark = pytest.mark.gpu
TOL_BF16 = 1e-2


def assert_close(a, b, tol, what, block=128):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, what
    g = rel_err(a, b)
    assert g < tol, "%s: relative L2 error %.3e >= %.1e" % (what, g, tol)
    emax = float(np.abs(a - b).max())
    bmax = float(np.abs(b).max())
    assert emax <= tol * bmax, "%s: max|a-b| = %.3e > %.1e * max|b| = %.3e" % (what, emax, tol, tol * bmax)
    if a.ndim == 4 and a.shape[2] * a.shape[3] >= block:
        # [N, C, H, W] -> per (image, block of `block` consecutive pixels): every channel of those pixels
        n, c, h, w = a.shape
        hw = (h * w) // block * block
        da = (a - b).reshape(n, c, h * w)[:, :, :hw].reshape(n, c, hw // block, block)
        db = b.reshape(n, c, h * w)[:, :, :hw].reshape(n, c, hw // block, block)
        num = np.sqrt((da ** 2).sum(axis=(1, 3)))
        den = np.sqrt((db ** 2).sum(axis=(1, 3)))
        floor = np.sqrt((b ** 2).mean()) * np.sqrt(c * block) * 1e-3   # blocks that are (almost) all zero
        worst = float((num / np.maximum(den, floor)).max())
        assert worst <= 3 * tol, "%s: worst %d-pixel block has relative error %.3e > %.1e" % (what, block, worst, 3 * tol)
    return g


def make_case(seed, N, C, H, W, O, modulated, sigma=2.0, k=3, stride=1, pad=1, dil=1, wscale=0.02):
    g = torch.Generator().manual_seed(seed)
    kh, kw = (k, k) if isinstance(k, int) else k
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    c = dict(x=torch.randn(N, C, H, W, generator=g), weight=torch.randn(O, C, kh, kw, generator=g) * wscale,
             offset=torch.randn(N, 2 * kh * kw, Ho, Wo, generator=g) * sigma,
             grad_out=torch.randn(N, O, Ho, Wo, generator=g), mask=None, bias=None,
             kw=dict(stride=stride, padding=pad, dilation=dil))
    if modulated:
        c["mask"] = torch.sigmoid(torch.randn(N, kh * kw, Ho, Wo, generator=g))
        c["bias"] = torch.randn(O, generator=g)
    return c


def np_(t):
    return None if t is None else t.numpy()


def run_gpu(c, dtype, backward=True, math="bf16"):
    dev = "cuda"
    x = c["x"].to(dev, dtype).requires_grad_(backward)
    w = c["weight"].to(dev, dtype).requires_grad_(backward)
    off = c["offset"].to(dev).requires_grad_(backward)
    m = None if c["mask"] is None else c["mask"].to(dev).requires_grad_(backward)
    b = None if c["bias"] is None else c["bias"].to(dev, dtype).requires_grad_(backward)
    kw = c["kw"]
    with sdb.dcn_math(math), torch.set_grad_enabled(backward):
        if m is None:
            y = sdb.deform_conv(x, off, w, kw["stride"], kw["padding"], kw["dilation"], 1, 1)
        else:
            y = sdb.modulated_deform_conv(x, off, m, w, b, kw["stride"], kw["padding"], kw["dilation"], 1, 1)
        if backward:
            y.backward(c["grad_out"].to(dev, dtype))
    torch.cuda.synchronize()
    f = lambda t: None if t is None else t.detach().float().cpu().numpy()
    out = dict(out=f(y))
    if backward:
        out.update(grad_x=f(x.grad), grad_offset=f(off.grad), grad_weight=f(w.grad),
                   grad_mask=None if m is None else f(m.grad), grad_bias=None if b is None else f(b.grad))
    return out


def oracle(c, backward=True, quantize=False):
    """CPU oracle.  quantize=True feeds it the bf16-rounded x / weight / grad_out the kernel is given when the
    TENSORS are bf16, so the comparison measures the kernel, not the input rounding."""
    q = (lambda t: t.bfloat16().float()) if quantize else (lambda t: t)
    x, w, gy = q(c["x"]), q(c["weight"]), q(c["grad_out"])
    b = None if c["bias"] is None else q(c["bias"])
    y = odcn.forward(np_(x), np_(c["offset"]), np_(w), mask=np_(c["mask"]), bias=np_(b), **c["kw"])
    res = dict(out=y)
    if backward:
        res.update(odcn.backward(np_(x), np_(c["offset"]), np_(w), np_(gy), mask=np_(c["mask"]),
                                 with_bias=b is not None, **c["kw"]))
    return res


def compare(got, ref, tol, tag):
    errs = {}
    for k, v in ref.items():
        if v is None:
            continue
        errs[k] = assert_close(got[k], v, tol, "%s %s" % (tag, k))
    return errs


# ---- BASELINE.json configs[1] / configs[0]: the benchmarked P3 map, and the reference test's 100x152 map --------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32io", "bf16io"])
@pytest.mark.parametrize("hw", [(100, 168), (100, 152)], ids=["P3_100x168_263tiles", "cfg0_100x152_238tiles"])
def test_benchmarked_p3_map_forward_and_all_gradients(hw, dtype):
    H, W = hw
    c = make_case(1000 + W, 2, 256, H, W, 256, False, sigma=2.0, wscale=0.01)
    got = run_gpu(c, dtype)
    ref = oracle(c, quantize=dtype == torch.bfloat16)
    errs = compare(got, ref, TOL_BF16, "N=2 256x%dx%d %s" % (H, W, dtype))
    print("rel errors", {k: "%.2e" % v for k, v in errs.items()})


def test_config2_modulated_batch8_forward_and_all_gradients():
    """BASELINE.json configs[2]: modulated DCN (mask + bias), batch 8 on the P3 map = 1 050 tiles (7.1 per CTA)."""
    c = make_case(2002, 8, 256, 100, 168, 256, True, sigma=2.0, wscale=0.01)
    got = run_gpu(c, torch.bfloat16)
    ref = oracle(c, quantize=True)
    errs = compare(got, ref, TOL_BF16, "configs[2] N=8 modulated")
    print("rel errors", {k: "%.2e" % v for k, v in errs.items()})


def test_config4_batch16_forward():
    """BASELINE.json configs[4] per-GPU batch: 16 x 256 x 100 x 168 = 2 100 tiles (14.2 per CTA)."""
    c = make_case(4004, 16, 256, 100, 168, 256, False, sigma=2.0, wscale=0.01)
    got = run_gpu(c, torch.bfloat16, backward=False)
    ref = oracle(c, backward=False, quantize=True)
    compare(got, ref, TOL_BF16, "configs[4] N=16 forward")


@pytest.mark.parametrize("hw", [(50, 84), (25, 42), (13, 21), (7, 11)], ids=["P4", "P5", "P6", "P7"])
def test_remaining_head_levels_bf16_io(hw):
    H, W = hw
    c = make_case(300 + H, 2, 256, H, W, 256, False, sigma=2.0, wscale=0.01)
    got = run_gpu(c, torch.bfloat16)
    ref = oracle(c, quantize=True)
    compare(got, ref, TOL_BF16, "N=2 256x%dx%d bf16" % (H, W))


def test_grad_input_run_to_run():
    """grad_input comes from a gather over the transposed sampling index; the index is put in a canonical order
    (entries of every list sorted by output pixel), so two runs on the same inputs must agree bit for bit."""
    c = make_case(77, 2, 256, 100, 168, 256, False, sigma=2.0, wscale=0.01)
    a = run_gpu(c, torch.bfloat16)
    b = run_gpu(c, torch.bfloat16)
    for k in ("out", "grad_x", "grad_offset", "grad_weight"):
        assert np.array_equal(a[k], b[k]), "%s differs between two runs (max %.3e)" % (k, np.abs(a[k] - b[k]).max())


# ---- tensor-core BACKWARD outside 3x3 / stride 1 / pad 1 (the forward cases of test_gpu_dcn.py, with gradients) ----
@pytest.mark.parametrize("shape", [
    # N, C, H, W, O, k, stride, pad, dil, modulated, sigma
    (2, 192, 19, 22, 48, 3, 2, 1, 1, False, 1.0),
    (2, 64, 33, 41, 64, 3, 2, 2, 2, True, 1.5),
    (1, 64, 12, 9, 16, 1, 1, 0, 1, False, 1.0),
    (1, 128, 30, 26, 128, 3, 1, 2, 2, True, 3.0),
    (1, 64, 11, 12, 32, (2, 4), 1, 1, 1, False, 1.0),
    (3, 256, 20, 19, 240, 3, 1, 1, 1, False, 30.0),
    (2, 64, 24, 30, 80, 3, 1, 0, 1, True, 2.0),
    (1, 512, 14, 13, 128, 3, 1, 1, 1, False, 2.0),
])
def test_tensor_core_backward_geometries(shape):
    N, C, H, W, O, k, st, pd, dl, mod, sigma = shape
    c = make_case(C + O + H, N, C, H, W, O, mod, sigma=sigma, k=k, stride=st, pad=pd, dil=dl, wscale=0.05)
    got = run_gpu(c, torch.float32)
    ref = oracle(c)
    compare(got, ref, TOL_BF16, "tc backward %r" % (shape,))


# ---- DFConv2d values (slender_det/layers/df_conv.py:6-79) ------------------------------------------
@pytest.mark.parametrize("modulated", [False, True])
def test_dfconv2d_values_and_gradients(modulated):
    """The module's output and every gradient against the same graph on the CPU: the offset(+mask) convolution in
    torch, the deformable convolution by the C oracle (autograd Function around it)."""
    class OracleDCN(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, off, m, w, b):
            ctx.save_for_backward(x, off, m, w)
            ctx.has = (m.numel() > 0, b.numel() > 0)
            y = odcn.forward(x.numpy(), off.numpy(), w.numpy(), mask=m.numpy() if ctx.has[0] else None,
                             bias=b.numpy() if ctx.has[1] else None, stride=1, padding=1)
            return torch.from_numpy(y)

        @staticmethod
        def backward(ctx, gy):
            x, off, m, w = ctx.saved_tensors
            g = odcn.backward(x.numpy(), off.numpy(), w.numpy(), gy.contiguous().numpy(),
                              mask=m.numpy() if ctx.has[0] else None, with_bias=ctx.has[1], stride=1, padding=1)
            t = lambda a, like: torch.from_numpy(a) if a is not None else torch.zeros_like(like)
            return (t(g["grad_x"], x), t(g["grad_offset"], off), t(g["grad_mask"], m), t(g["grad_weight"], w),
                    t(g["grad_bias"], gy[0, :, 0, 0]))

    torch.backends.cudnn.allow_tf32 = False   # the module's plain offset conv must be exact fp32 for a 1e-4 comparison
    torch.manual_seed(5)
    mod = sdb.DFConv2d(64, 64, with_modulated_dcn=modulated, bias=modulated)
    with torch.no_grad():
        mod.offset.weight.mul_(0.3)
        mod.offset.bias.uniform_(-1.0, 1.0)
    x = torch.randn(2, 64, 21, 27)
    gy = torch.randn(2, 64, 21, 27)
    # CPU graph
    xc = x.clone().requires_grad_()
    pred = torch.nn.functional.conv2d(xc, mod.offset.weight, mod.offset.bias, padding=1)
    if modulated:
        off, m = pred[:, :18].contiguous(), pred[:, 18:].sigmoid().contiguous()
    else:
        off, m = pred.contiguous(), torch.zeros(0)
    bias = mod.conv.bias if mod.conv.bias is not None else torch.zeros(0)
    yc = OracleDCN.apply(xc, off, m, mod.conv.weight, bias)
    yc.backward(gy)
    ref = dict(out=yc.detach().numpy(), grad_x=xc.grad.numpy())
    for n, p in mod.named_parameters():
        ref["grad_" + n] = p.grad.numpy().copy()
        p.grad = None
    # GPU module, exact fp32 kernels
    mg = mod.cuda()
    xg = x.cuda().requires_grad_()
    with sdb.dcn_math("fp32"):
        yg = mg(xg)
        yg.backward(gy.cuda())
    torch.cuda.synchronize()
    assert rel_err(yg.detach().cpu().numpy(), ref["out"]) < 1e-4
    assert rel_err(xg.grad.cpu().numpy(), ref["grad_x"]) < 1e-4
    for n, p in mg.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), ref["grad_" + n]) < 2e-4, n
