"""GPU parity of the whole-head calls (sdb_dcn_forward_multi / sdb_dcn_backward_multi through
slenderobjdet_b200.deform_conv_multi): every FPN level x every convolution in ONE launch per kernel, operand images of
the weights prepared once, one transposed sampling index per offset group.

Reference loop being replaced: slender_det/modeling/meta_arch/reppoints/reppointsv2.py:728-752 (per level: two
DeformConv calls that consume the same dcn_offset).  Checked against the CPU oracle (which is pinned on the reference,
tests/test_oracle_dcn.py, tests/test_gpu_ref_c.py) and against the per-problem calls.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
import slenderobjdet_b200 as sdb
from oracle import dcn as odcn
from test_gpu_dcn_large import assert_close

pytestmark = pytest.mark.gpu
LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]   # P3..P7 of an 800x1344 image


def _head_case(seed, N, C, O, levels, modulated=False, sigma=2.0):
    g = torch.Generator().manual_seed(seed)
    c = dict(w=[torch.randn(O, C, 3, 3, generator=g) * 0.01 for _ in range(2)], lv=[])
    if modulated:
        c["b"] = [torch.randn(O, generator=g) for _ in range(2)]
    for (H, W) in levels:
        lv = dict(off=torch.randn(N, 18, H, W, generator=g) * sigma,
                  x=[torch.randn(N, C, H, W, generator=g) for _ in range(2)],
                  gy=[torch.randn(N, O, H, W, generator=g) for _ in range(2)])
        if modulated:
            lv["m"] = torch.sigmoid(torch.randn(N, 9, H, W, generator=g))
        c["lv"].append(lv)
    return c


def _run_multi(c, dtype, math="bf16"):
    """The head as one call per pass: problems ordered (level, branch); both branches of a level get the SAME offset
    tensor object, as in the reference's forward."""
    dev = "cuda"
    ws = [w.to(dev, dtype).requires_grad_() for w in c["w"]]
    bs = [b.to(dev, dtype).requires_grad_() for b in c["b"]] if "b" in c else None
    xs, offs, masks, wids, gys, off_leaf, mask_leaf = [], [], [], [], [], [], []
    for lv in c["lv"]:
        off = lv["off"].to(dev).requires_grad_()
        m = lv["m"].to(dev).requires_grad_() if "m" in lv else None
        off_leaf.append(off)
        mask_leaf.append(m)
        for b in range(2):
            xs.append(lv["x"][b].to(dev, dtype).requires_grad_())
            offs.append(off)
            masks.append(m)
            wids.append(b)
            gys.append(lv["gy"][b].to(dev, dtype))
    with sdb.dcn_math(math):
        ys = sdb.deform_conv_multi(xs, offs, ws, 1, 1, 1, masks=masks if "b" in c else None, biases=bs, weight_ids=wids)
        torch.autograd.backward(ys, gys)
    torch.cuda.synchronize()
    f = lambda t: t.detach().float().cpu().numpy()
    return dict(out=[f(y) for y in ys], gx=[f(x.grad) for x in xs], goff=[f(o.grad) for o in off_leaf],
                gmask=[f(m.grad) for m in mask_leaf] if "b" in c else None, gw=[f(w.grad) for w in ws],
                gb=[f(b.grad) for b in bs] if bs else None)


def _oracle_head(c, quantize):
    q = (lambda t: t.bfloat16().float().numpy()) if quantize else (lambda t: t.numpy())
    res = dict(out=[], gx=[], goff=[], gmask=[], gw=[np.zeros_like(c["w"][0].numpy(), dtype=np.float64) for _ in range(2)],
               gb=[np.zeros(c["w"][0].shape[0], np.float64) for _ in range(2)])
    for lv in c["lv"]:
        goff = 0.0
        gmask = 0.0
        for b in range(2):
            m = lv["m"].numpy() if "m" in lv else None
            bias = q(c["b"][b]) if "b" in c else None
            y = odcn.forward(q(lv["x"][b]), lv["off"].numpy(), q(c["w"][b]), mask=m, bias=bias, stride=1, padding=1)
            gr = odcn.backward(q(lv["x"][b]), lv["off"].numpy(), q(c["w"][b]), q(lv["gy"][b]), mask=m,
                               with_bias=bias is not None, stride=1, padding=1)
            res["out"].append(y)
            res["gx"].append(gr["grad_x"])
            goff = goff + gr["grad_offset"].astype(np.float64)
            if m is not None:
                gmask = gmask + gr["grad_mask"].astype(np.float64)
                res["gb"][b] += gr["grad_bias"]
            res["gw"][b] += gr["grad_weight"]
        res["goff"].append(goff)      # the shared offset tensor receives the sum of both branches' gradients
        res["gmask"].append(gmask)
    return res


def _compare(got, ref, tol, modulated):
    for i, (a, b) in enumerate(zip(got["out"], ref["out"])):
        assert_close(a, b, tol, "out[%d]" % i)
    for i, (a, b) in enumerate(zip(got["gx"], ref["gx"])):
        assert_close(a, b, tol, "grad_x[%d]" % i)
    for i, (a, b) in enumerate(zip(got["goff"], ref["goff"])):
        assert_close(a, b, tol, "grad_offset[level %d]" % i)
    for i, (a, b) in enumerate(zip(got["gw"], ref["gw"])):
        assert_close(a, b, tol, "grad_weight[%d]" % i)
    if modulated:
        for i, (a, b) in enumerate(zip(got["gmask"], ref["gmask"])):
            assert_close(a, b, tol, "grad_mask[level %d]" % i)
        for i, (a, b) in enumerate(zip(got["gb"], ref["gb"])):
            assert_close(a, b, tol, "grad_bias[%d]" % i)


def test_reppoints_head_pair_all_levels_vs_oracle():
    """BASELINE.json configs[1] exactly: both head DCNs (256 -> 256, 3x3) on P3-P7 of an 800x1344 image, batch 2, bf16
    tensors: 10 problems, 708 tiles, one launch per kernel.  Outputs and every gradient against the oracle, with the
    element-wise and per-block bounds."""
    c = _head_case(11, 2, 256, 256, LEVELS)
    got = _run_multi(c, torch.bfloat16)
    ref = _oracle_head(c, quantize=True)
    _compare(got, ref, 1e-2, False)


def test_modulated_head_with_bias_vs_oracle():
    c = _head_case(12, 2, 128, 128, LEVELS[1:], modulated=True)
    got = _run_multi(c, torch.bfloat16)
    ref = _oracle_head(c, quantize=True)
    _compare(got, ref, 1e-2, True)


def test_fp32_math_multi_vs_oracle():
    c = _head_case(13, 1, 64, 64, LEVELS[2:], modulated=True)
    got = _run_multi(c, torch.float32, math="fp32")
    ref = _oracle_head(c, quantize=False)
    _compare(got, ref, 1e-4, True)


def test_multi_matches_per_problem_calls():
    """Same kernels, same tiles: the outputs, grad_input and grad_offset of the one-launch path equal the per-problem
    calls bit for bit (the transposed index is in canonical order); grad_weight differs only by the split-K order."""
    c = _head_case(14, 2, 256, 256, LEVELS[1:])
    got = _run_multi(c, torch.bfloat16)
    i = 0
    gw = [0.0, 0.0]
    for li, lv in enumerate(c["lv"]):
        goff = 0.0
        for b in range(2):
            x = lv["x"][b].cuda().bfloat16().requires_grad_()
            off = lv["off"].cuda().requires_grad_()
            w = c["w"][b].cuda().bfloat16().requires_grad_()
            y = sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
            y.backward(lv["gy"][b].cuda().bfloat16())
            assert np.array_equal(got["out"][i], y.detach().float().cpu().numpy()), ("out", li, b)
            assert np.array_equal(got["gx"][i], x.grad.float().cpu().numpy()), ("grad_x", li, b)
            goff = goff + off.grad
            gw[b] = gw[b] + w.grad.float()
            i += 1
        assert rel_err(got["goff"][li], goff.cpu().numpy()) < 1e-6
    for b in range(2):
        # per-problem calls round each level's dW to bf16 before summing; the multi path sums in fp32
        assert rel_err(got["gw"][b], gw[b].cpu().numpy()) < 4e-3


def test_ragged_batches_single_weight_and_module_api():
    """Levels with different batch sizes, one weight, through DeformConv.forward_multi; also a problem with N = 0."""
    torch.manual_seed(3)
    conv = sdb.DeformConv(64, 64, 3, 1, 1).cuda()
    shapes = [(3, 20, 19), (1, 13, 21), (2, 7, 11)]
    xs = [torch.randn(n, 64, h, w, device="cuda", requires_grad=True) for n, h, w in shapes]
    offs = [torch.randn(n, 18, h, w, device="cuda", requires_grad=True) for n, h, w in shapes]
    with sdb.dcn_math("bf16"):
        ys = conv.forward_multi(xs, offs)
        sum(y.square().sum() for y in ys).backward()
    gw_multi = conv.weight.grad.clone()
    conv.weight.grad = None
    ref = []
    for x, o in zip(xs, offs):
        x2, o2 = x.detach().clone().requires_grad_(), o.detach().clone().requires_grad_()
        with sdb.dcn_math("bf16"):
            y = conv(x2, o2)
            y.square().sum().backward()
        ref.append((y, x2.grad, o2.grad))
    for (y, gx, go), ym, x, o in zip(ref, ys, xs, offs):
        assert torch.equal(y, ym) and torch.equal(gx, x.grad) and torch.equal(go, o.grad)
    assert rel_err(gw_multi.cpu().numpy(), conv.weight.grad.cpu().numpy()) < 1e-5


def test_prepared_weights_follow_weight_updates():
    """The operand images are cached per weight version: an in-place update (an optimiser step) must be seen."""
    torch.manual_seed(4)
    conv = sdb.DeformConv(64, 64, 3, 1, 1).cuda()
    x = torch.randn(1, 64, 12, 14, device="cuda")
    off = torch.randn(1, 18, 12, 14, device="cuda")
    with sdb.dcn_math("bf16"), torch.no_grad():
        y0 = conv(x, off)
        conv.weight.mul_(2.0)                 # bumps the version counter, like optimizer.step()
        y1 = conv(x, off)
        conv.weight.data = conv.weight.data * 0.5   # new storage
        y2 = conv(x, off)
    assert rel_err(y1.cpu().numpy(), (2 * y0).cpu().numpy()) < 1e-3
    assert rel_err(y2.cpu().numpy(), y0.cpu().numpy()) < 1e-3


def test_phased_backward_through_the_c_abi_equals_one_call():
    """sdb_dcn_backward_multi in phases -- the order bench.py uses to run the all-reduce beside the grad_input gather:
    DATA_ONLY|NO_GATHER, then WEIGHT_ONLY|GRAD_PACKED, then DATA_ONLY|GATHER_ONLY -- gives the bits of the one-call
    backward (same kernels, same workspace)."""
    import bench
    from slenderobjdet_b200 import _lib as L
    dev = torch.device("cuda", 0)
    wl = bench.Workload(torch, L, dev, seed=3, batch=2, levels=[(25, 42), (13, 21), (7, 11)])
    st = torch.cuda.current_stream(dev)

    brs = [br for lv in wl.lv for br in lv["br"]]

    def grads():
        return [b["gx"].clone() for b in brs] + [b["goff"].clone() for b in brs] + [g.clone() for g in wl.gw]

    def clear():
        for b in brs:
            b["gx"].zero_()
            b["goff"].zero_()
        for g in wl.gw:
            g.zero_()

    wl.phase_forward(st)
    clear()
    wl.phase_backward(st, 0)
    torch.cuda.synchronize()
    one = grads()
    clear()
    wl.phase_backward(st, L.SDB_BWD_DATA_ONLY | L.SDB_BWD_NO_GATHER)
    torch.cuda.synchronize()
    assert all(float(b["gx"].float().abs().max()) == 0.0 for b in brs)     # the gather has not run yet
    wl.phase_backward(st, L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_GRAD_PACKED)
    wl.phase_backward(st, L.SDB_BWD_DATA_ONLY | L.SDB_BWD_GATHER_ONLY | L.SDB_BWD_GRAD_PACKED)
    torch.cuda.synchronize()
    for a, b in zip(one, grads()):
        assert torch.equal(a, b)
    # the two-half order of bench.py at N > 1: the transposed index is built in the weight-gradient half
    clear()
    wl.phase_backward(st, L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_BUILD_INDEX)
    torch.cuda.synchronize()
    assert all(float(b["gx"].float().abs().max()) == 0.0 and float(b["goff"].abs().max()) == 0.0 for b in brs)
    wl.phase_backward(st, L.SDB_BWD_DATA_ONLY | L.SDB_BWD_GRAD_PACKED | L.SDB_BWD_INDEX_READY)
    torch.cuda.synchronize()
    for a, b in zip(one, grads()):
        assert torch.equal(a, b)


def test_cta_pair_kernels_give_the_bits_of_the_single_cta_kernels():
    """sdb_set_forward_pair / sdb_set_backward_pair only change how tiles are mapped to CTAs (cta_group::2, M = 256, two
    tiles per cluster, an all-invalid second tile when a problem's tile count is odd): every output and gradient is
    bit-identical to the one-CTA-per-tile kernels (include/slender_b200.h says so)."""
    import bench
    from slenderobjdet_b200 import _lib as L
    lib = L.lib()
    dev = torch.device("cuda", 0)
    wl = bench.Workload(torch, L, dev, seed=5, batch=1, levels=[(50, 84), (25, 42), (13, 21), (7, 11)])   # 33, 9, 3, 1 tiles
    st = torch.cuda.current_stream(dev)
    brs = [br for lv in wl.lv for br in lv["br"]]
    res = {}
    try:
        for pair in (0, 1):
            lib.sdb_set_forward_pair(pair)
            lib.sdb_set_backward_pair(pair)
            for b in brs:
                b["out"].zero_(); b["gx"].zero_(); b["goff"].zero_()
            wl.step(st)
            torch.cuda.synchronize()
            res[pair] = [b[k].clone() for b in brs for k in ("out", "gx", "goff")] + [g.clone() for g in wl.gw]
    finally:
        lib.sdb_set_forward_pair(1)
        lib.sdb_set_backward_pair(1)
    assert all(torch.equal(a, b) for a, b in zip(res[0], res[1]))
    assert float(res[1][0].float().abs().max()) > 0
