"""GPU parity of the tower path: the plain 'same' convolution as the zero-offset specialisation of
the tcgen05 deformable-convolution kernels (offset == NULL through the C ABI), and whole towers
Conv2d -> GroupNorm -> ReLU (reppointsv2.py:644-675, :733-736).

Oracles: oracle/dcn_oracle.c with ZERO offsets (a deformable convolution with zero offsets is the convolution,
tests/test_deformable_conv.py:85-87 of the reference asserts exactly that) for the convolution, and the reference's own
layers -- torch.nn.Conv2d / GroupNorm / ReLU in float64 on the CPU -- for whole towers.  bf16 arithmetic: rel <= 1e-2
with the element-wise and per-128-pixel-block bounds of test_gpu_dcn_large.assert_close."""
import numpy as np
import pytest
import torch

import slenderobjdet_b200 as sdb
import slenderobjdet_b200.layers as L
from oracle import dcn as odcn
from test_gpu_dcn_large import assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-2
bf = torch.bfloat16


def _conv_case(seed, N, C, H, W, O, k=3, pad=1, dil=1, bias=False):
    g = torch.Generator().manual_seed(seed)
    q = lambda t: t.to(bf)
    c = dict(x=q(torch.randn(N, C, H, W, generator=g)), w=q(torch.randn(O, C, k, k, generator=g) * 0.03),
             gy=q(torch.randn(N, O, H, W, generator=g)), b=q(torch.randn(O, generator=g)) if bias else None,
             kw=dict(stride=1, padding=pad, dilation=dil))
    return c


def _oracle_conv(c):
    x, w, gy = c["x"].float().numpy(), c["w"].float().numpy(), c["gy"].float().numpy()
    N, _, H, W = x.shape
    k = w.shape[2]
    off = np.zeros((N, 2 * k * k, H, W), np.float32)
    b = None if c["b"] is None else c["b"].float().numpy()
    y = odcn.forward(x, off, w, bias=b, **c["kw"])
    gr = odcn.backward(x, off, w, gy, with_bias=b is not None, **c["kw"])
    return y, gr


SHAPES = [(2, 256, 25, 42, 256, 3, 1, 1), (1, 64, 13, 21, 128, 3, 1, 1), (2, 128, 20, 19, 64, 3, 2, 2),
          (2, 64, 9, 11, 64, 1, 0, 1), (3, 192, 7, 11, 192, 3, 1, 1)]


@pytest.mark.parametrize("bias", [False, True], ids=["nobias", "bias"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "N%dC%d_%dx%d_O%d_k%dp%dd%d" % s)
def test_plain_convolution_forward_and_gradients(shape, bias):
    N, C, H, W, O, k, pad, dil = shape
    c = _conv_case(C + O + H, N, C, H, W, O, k, pad, dil, bias)
    x = c["x"].cuda().requires_grad_()
    w = c["w"].cuda().requires_grad_()
    b = c["b"].cuda().requires_grad_() if bias else None
    y = L.conv2d_multi([x], [w], [b] if bias else None, pad, dil)[0]
    y.backward(c["gy"].cuda())
    torch.cuda.synchronize()
    yo, go = _oracle_conv(c)
    f = lambda t: t.detach().float().cpu().numpy()
    assert_close(f(y), yo, TOL, "out")
    assert_close(f(x.grad), go["grad_x"], TOL, "grad_x")
    assert_close(f(w.grad), go["grad_weight"], TOL, "grad_weight")
    if bias:
        assert_close(f(b.grad), go["grad_bias"], TOL, "grad_bias")
    # and against the reference's own layer (torch conv2d, float64 on the CPU)
    yt = torch.nn.functional.conv2d(c["x"].double(), c["w"].double(), None if not bias else c["b"].double(), 1, pad, dil)
    assert_close(f(y), yt.numpy(), TOL, "out vs torch conv2d")


def test_plain_convolution_benchmarked_p3_map():
    """2 x 256 x 100 x 168 = 263 tiles on 148 persistent CTAs, forward and both gradients"""
    c = _conv_case(7, 2, 256, 100, 168, 256)
    x, w = c["x"].cuda().requires_grad_(), c["w"].cuda().requires_grad_()
    y = L.conv2d_multi([x], [w])[0]
    y.backward(c["gy"].cuda())
    torch.cuda.synchronize()
    yo, go = _oracle_conv(c)
    f = lambda t: t.detach().float().cpu().numpy()
    assert_close(f(y), yo, TOL, "out")
    assert_close(f(x.grad), go["grad_x"], TOL, "grad_x")
    assert_close(f(w.grad), go["grad_weight"], TOL, "grad_weight")


def test_float32_tensors_need_an_explicit_opt_in():
    x = torch.randn(1, 64, 8, 8, device="cuda")
    w = torch.randn(64, 64, 3, 3, device="cuda")
    with pytest.raises(RuntimeError, match="not silently demoted"):
        L.conv2d_multi([x], [w])
    with sdb.dcn_math("bf16"):
        y = L.conv2d_multi([x], [w])[0]
    ref = torch.nn.functional.conv2d(x.double().cpu(), w.double().cpu(), padding=1)
    assert y.dtype == torch.float32
    assert_close(y.cpu().numpy(), ref.numpy(), TOL, "fp32 tensors, bf16 math")
    with pytest.raises(RuntimeError, match="same"):
        with sdb.dcn_math("bf16"):
            L.conv2d_multi([x], [w], padding=0)


def test_towers_match_the_reference_layers():
    """two towers x three FPN levels x two (conv, GN, ReLU) layers: outputs, input gradient and every parameter gradient
    against nn.Conv2d / nn.GroupNorm / nn.ReLU(inplace) in float64; state-dict keys as in the reference"""
    torch.manual_seed(11)
    C, depth = 64, 2
    levels = [(20, 34), (10, 17), (5, 9)]
    towers = [L.build_tower(C, C, depth).cuda().to(bf) for _ in range(2)]
    refs = []
    for tw in towers:
        seq = torch.nn.ModuleList()
        for i in range(depth):
            seq.append(torch.nn.Conv2d(C, C, 3, 1, 1, bias=False))
            seq.append(torch.nn.GroupNorm(32 * C // 256, C))
            seq.append(torch.nn.ReLU(inplace=True))
        with torch.no_grad():
            for m in seq:
                if isinstance(m, torch.nn.Conv2d):
                    m.weight.normal_(0, 0.05)
                if isinstance(m, torch.nn.GroupNorm):
                    m.weight.normal_(1, 0.2)
                    m.bias.normal_(0, 0.2)
        sd = {k: v.to(bf) for k, v in seq.state_dict().items()}
        tw.load_state_dict(sd)                       # the reference's keys load as they are
        seq.load_state_dict({k: v.float() for k, v in sd.items()})
        refs.append(seq.double())
    feats = [torch.randn(2, C, H, W).to(bf) for (H, W) in levels]
    gys = [[torch.randn(2, C, H, W).to(bf) for (H, W) in levels] for _ in range(2)]
    fd = [f.cuda().requires_grad_() for f in feats]
    outs = L.towers_forward(towers, fd)
    torch.autograd.backward([o for t in outs for o in t], [g.cuda() for t in gys for g in t])
    torch.cuda.synchronize()
    fr = [f.double().requires_grad_() for f in feats]
    # the kernels keep activations in bf16 between layers: the reference rounds where they do (straight-through for the
    # gradient), so the comparison measures the kernels' arithmetic and not ReLU masks flipped by activation rounding
    rnd = lambda t: t + (t.detach().to(bf).double() - t.detach())
    routs = []
    for seq in refs:
        lv = []
        for f in fr:
            h = f
            for m in seq:
                h = m(h)
                if not isinstance(m, torch.nn.GroupNorm):   # conv output and the fused GroupNorm+ReLU output are stored
                    h = rnd(h)
            lv.append(h)
        routs.append(lv)
    torch.autograd.backward([o for t in routs for o in t], [g.double() for t in gys for g in t])
    # two bf16 layers deep: activations are re-rounded to bf16 between layers, so the bound is 2x the per-op tolerance
    tol = 2 * TOL
    f = lambda t: t.detach().float().cpu().numpy()
    for t in range(2):
        for l in range(len(levels)):
            assert_close(f(outs[t][l]), routs[t][l].detach().numpy(), tol, "tower %d level %d out" % (t, l))
    for l in range(len(levels)):
        assert_close(f(fd[l].grad), fr[l].grad.numpy(), tol, "level %d grad_input" % l)
    for t in range(2):
        for (k, p), (_, q) in zip(towers[t].named_parameters(), refs[t].named_parameters()):
            assert_close(f(p.grad), q.grad.numpy(), tol, "tower %d %s grad" % (t, k))
