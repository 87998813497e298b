"""GPU parity of the reference-granularity C entry points called DIRECTLY through ctypes (the calls a
`detectron2._C`-style binding would make, include/slender_b200.h): sdb_dcn_forward, sdb_dcn_backward_data (which
ACCUMULATES into grad_x, deform_conv.py:89-90) and sdb_dcn_backward_weight (which accumulates into grad_weight /
grad_bias with `scale`, deform_conv_cuda.cu:770-777), plus selective gradients through the whole-head call."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import rel_err
import slenderobjdet_b200 as sdb
from slenderobjdet_b200 import _lib as L
from oracle import dcn as odcn

pytestmark = pytest.mark.gpu


def _case(seed, N, C, H, W, O, modulated):
    g = torch.Generator().manual_seed(seed)
    c = dict(x=torch.randn(N, C, H, W, generator=g), w=torch.randn(O, C, 3, 3, generator=g) * 0.05,
             off=torch.randn(N, 18, H, W, generator=g) * 2.0, gy=torch.randn(N, O, H, W, generator=g))
    if modulated:
        c["m"] = torch.sigmoid(torch.randn(N, 9, H, W, generator=g))
        c["b"] = torch.randn(O, generator=g)
    return c


@pytest.mark.parametrize("dtype,math", [(torch.bfloat16, "bf16"), (torch.float32, "bf16"), (torch.float32, "fp32")])
@pytest.mark.parametrize("modulated", [False, True])
def test_single_problem_entry_points(dtype, math, modulated):
    N, C, H, W, O = 2, 128, 19, 22, 80
    c = _case(11 + modulated, N, C, H, W, O, modulated)
    lib = L.lib()
    dev = "cuda"
    mth = L.SDB_MATH_BF16 if math == "bf16" else L.SDB_MATH_FP32
    iod = L.SDB_BF16 if dtype == torch.bfloat16 else L.SDB_F32
    geom = L.Geom(N, C, H, W, O, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    gp = ctypes.byref(geom)
    x, w, gy = c["x"].to(dev, dtype), c["w"].to(dev, dtype), c["gy"].to(dev, dtype)
    off = c["off"].to(dev)
    m = c["m"].to(dev) if modulated else None
    b = c["b"].to(dev, dtype) if modulated else None
    q = (lambda t: t.to(dtype).float().numpy()) if math == "bf16" else (lambda t: t.numpy())
    mo = c["m"].numpy() if modulated else None
    yo = odcn.forward(q(c["x"]), c["off"].numpy(), q(c["w"]), mask=mo, bias=q(c["b"]) if modulated else None, stride=1, padding=1)
    ref = odcn.backward(q(c["x"]), c["off"].numpy(), q(c["w"]), q(c["gy"]), mask=mo, with_bias=modulated, stride=1, padding=1)
    tol = 1e-2 if math == "bf16" else 1e-4
    st = L.stream_ptr(torch.device(dev))

    def ws(op):
        n = int(lib.sdb_dcn_workspace_bytes(op, gp, iod, mth))
        return torch.empty(max(n, 1), dtype=torch.uint8, device=dev), n

    # forward (exporting the packed input for the two backward calls)
    out = torch.empty(N, O, H, W, device=dev, dtype=dtype)
    pkb = int(lib.sdb_dcn_packed_input_bytes(gp, mth))
    pk = torch.empty(max(pkb, 1), dtype=torch.uint8, device=dev)
    wsf, nf = ws(L.SDB_OP_FORWARD)
    L.check(lib.sdb_dcn_forward(L.ptr(x), L.ptr(off), L.ptr(m), L.ptr(w), L.ptr(b), L.ptr(out), gp, iod, mth, L.ptr(wsf), nf,
                                L.ptr(pk) if pkb else None, st))
    assert rel_err(out.float().cpu().numpy(), yo) < tol

    # backward_data: grad_x is ACCUMULATED into, grad_offset / grad_mask are overwritten
    base = torch.randn(N, C, H, W, device=dev).to(dtype)
    gx = base.clone()
    goff = torch.full((N, 18, H, W), 7.0, device=dev)
    gm = torch.full((N, 9, H, W), 7.0, device=dev) if modulated else None
    wsd, nd = ws(L.SDB_OP_BACKWARD_DATA)
    L.check(lib.sdb_dcn_backward_data(L.ptr(x), L.ptr(off), L.ptr(m), L.ptr(w), L.ptr(gy), L.ptr(gx), L.ptr(goff), L.ptr(gm),
                                      gp, iod, mth, L.ptr(wsd), nd, L.ptr(pk) if pkb else None, st))
    got_gx = (gx.float() - base.float()).cpu().numpy()
    # bf16 tensors: the accumulated sum is rounded once more, to the magnitude of base + grad
    assert rel_err(got_gx, ref["grad_x"]) < (2e-2 if dtype == torch.bfloat16 else tol)
    assert rel_err(goff.cpu().numpy(), ref["grad_offset"]) < tol
    if modulated:
        assert rel_err(gm.cpu().numpy(), ref["grad_mask"]) < tol

    # backward_weight: grad_weight += scale * dW, grad_bias += scale * sum dY (float32 always)
    gw = torch.ones(O, C, 3, 3, device=dev)
    gb = torch.ones(O, device=dev) if modulated else None
    wsw, nw = ws(L.SDB_OP_BACKWARD_WEIGHT)
    L.check(lib.sdb_dcn_backward_weight(L.ptr(x), L.ptr(off), L.ptr(m), L.ptr(gy), L.ptr(gw), L.ptr(gb), ctypes.c_float(0.5), gp,
                                        iod, mth, L.ptr(wsw), nw, L.ptr(pk) if pkb else None, st))
    torch.cuda.synchronize()
    assert rel_err((gw.cpu().numpy() - 1.0) / 0.5, ref["grad_weight"]) < tol
    if modulated:
        assert rel_err((gb.cpu().numpy() - 1.0) / 0.5, ref["grad_bias"]) < tol


@pytest.mark.parametrize("which", ["x", "offset", "weight", "x+weight"])
def test_selective_gradients_whole_head_call(which):
    """Only some inputs require grad: the backward must skip the other kernels (the grad_offset kernel then only exports
    dcol for the gather, or the gather / weight gradient do not run at all) and still be right."""
    N, C, O = 2, 128, 64
    levels = [(13, 21), (7, 11)]
    g = torch.Generator().manual_seed(3)
    w = torch.randn(O, C, 3, 3, generator=g) * 0.05
    xs = [torch.randn(N, C, h, ww, generator=g) for h, ww in levels]
    offs = [torch.randn(N, 18, h, ww, generator=g) * 2 for h, ww in levels]
    gys = [torch.randn(N, O, h, ww, generator=g) for h, ww in levels]
    wd = w.cuda().bfloat16().requires_grad_("weight" in which)
    xd = [x.cuda().bfloat16().requires_grad_("x" in which) for x in xs]
    od = [o.cuda().requires_grad_("offset" in which) for o in offs]
    with sdb.dcn_math("bf16"):
        ys = sdb.deform_conv_multi(xd, od, wd, 1, 1, 1)
        torch.autograd.backward(ys, [t.cuda().bfloat16() for t in gys])
    torch.cuda.synchronize()
    q = lambda t: t.bfloat16().float().numpy()
    gw_ref = 0.0
    for i in range(len(levels)):
        ref = odcn.backward(q(xs[i]), offs[i].numpy(), q(w), q(gys[i]), stride=1, padding=1)
        gw_ref = gw_ref + ref["grad_weight"].astype(np.float64)
        if "x" in which:
            assert rel_err(xd[i].grad.float().cpu().numpy(), ref["grad_x"]) < 1e-2
        else:
            assert xd[i].grad is None
        if "offset" in which:
            assert rel_err(od[i].grad.cpu().numpy(), ref["grad_offset"]) < 1e-2
        else:
            assert od[i].grad is None
    if "weight" in which:
        assert rel_err(wd.grad.float().cpu().numpy(), gw_ref) < 1e-2
    else:
        assert wd.grad is None
