"""GPU parity of the fused GroupNorm + ReLU kernels (csrc/gn.cu, through the Python operator API -> ctypes -> C ABI)
against the float64 oracle (oracle/gn.py, pinned on torch.nn.GroupNorm + nn.ReLU).  Tolerances: float32 tensors rel
<= 1e-5 (statistics are combined in fp64), bfloat16 tensors rel <= 1e-2 against the oracle fed the same bf16-rounded
inputs (output rounding only)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
import slenderobjdet_b200.layers as L
from oracle import gn as ogn

pytestmark = pytest.mark.gpu
LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]


def _case(seed, N, C, H, W, dtype):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(N, C, H, W, generator=g) * 1.5 + 0.25).to(dtype)
    gy = torch.randn(N, C, H, W, generator=g).to(dtype)
    return x, gy


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", [(2, 256, 25, 42, 32), (1, 64, 7, 11, 8), (3, 32, 5, 3, 32), (2, 256, 100, 168, 32)],
                         ids=lambda s: "N%dC%d_%dx%d_G%d" % s)
def test_group_norm_relu_vs_oracle(shape, relu, dtype):
    N, C, H, W, G = shape
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    x, gy = _case(C + H, N, C, H, W, dtype)
    g = torch.Generator().manual_seed(1)
    gamma = torch.randn(C, generator=g) * 0.5 + 1.0
    beta = torch.randn(C, generator=g) * 0.2
    xd = x.cuda().requires_grad_()
    gd, bd = gamma.cuda().requires_grad_(), beta.cuda().requires_grad_()
    y = L.group_norm_relu(xd, G, gd, bd, 1e-5, relu)
    y.backward(gy.cuda())
    torch.cuda.synchronize()
    xf, gyf = x.float().numpy(), gy.float().numpy()
    yo, _, _ = ogn.forward(xf, gamma.numpy(), beta.numpy(), G, 1e-5, relu)
    go = ogn.backward(xf, gamma.numpy(), beta.numpy(), G, gyf, 1e-5, relu)
    assert y.dtype == dtype and xd.grad.dtype == dtype
    assert rel_err(y.detach().float().cpu().numpy(), yo) < tol
    assert rel_err(xd.grad.float().cpu().numpy(), go["grad_x"]) < tol
    assert rel_err(gd.grad.cpu().numpy(), go["grad_gamma"]) < max(tol, 2e-5) * (1 if dtype == torch.float32 else 0.3)
    assert rel_err(bd.grad.cpu().numpy(), go["grad_beta"]) < max(tol, 2e-5) * (1 if dtype == torch.float32 else 0.3)
    if dtype == torch.float32:   # element-wise, too
        assert np.abs(y.detach().cpu().numpy() - yo).max() <= 1e-5 * max(np.abs(yo).max(), 1.0)


def test_multi_call_over_levels_and_towers_equals_single_calls_and_is_reproducible():
    """one call over 5 FPN levels x 2 towers (two parameter sets): same bits as ten single calls; parameter gradients
    accumulate over the levels in a fixed order, so two runs agree bit for bit"""
    g = torch.Generator().manual_seed(3)
    C, G, N = 64, 8, 2
    gam = [(torch.randn(C, generator=g) * 0.3 + 1).cuda().requires_grad_() for _ in range(2)]
    bet = [(torch.randn(C, generator=g) * 0.3).cuda().requires_grad_() for _ in range(2)]
    xs, gys, pids = [], [], []
    for (H, W) in LEVELS:
        for k in range(2):
            xs.append(torch.randn(N, C, H // 4 + 1, W // 4 + 1, generator=g).cuda().requires_grad_())
            gys.append(torch.randn(N, C, H // 4 + 1, W // 4 + 1, generator=g).cuda())
            pids.append(k)
    runs = []
    for _ in range(2):
        for t in xs + gam + bet:
            t.grad = None
        ys = L.group_norm_relu_multi(xs, gam, bet, G, 1e-5, pids)
        torch.autograd.backward(ys, gys)
        runs.append([y.detach().clone() for y in ys] + [x.grad.clone() for x in xs] + [t.grad.clone() for t in gam + bet])
    assert all(torch.equal(a, b) for a, b in zip(*runs))
    gsum = [torch.zeros(C, device="cuda", dtype=torch.float64) for _ in range(4)]
    for i, x in enumerate(xs):
        x1 = x.detach().clone().requires_grad_()
        g1, b1 = gam[pids[i]].detach().clone().requires_grad_(), bet[pids[i]].detach().clone().requires_grad_()
        y1 = L.group_norm_relu(x1, G, g1, b1)
        y1.backward(gys[i])
        assert torch.equal(y1, runs[0][i]) and torch.equal(x1.grad, runs[0][len(xs) + i])
        gsum[pids[i]] += g1.grad.double()
        gsum[2 + pids[i]] += b1.grad.double()
    for k in range(4):
        assert rel_err(runs[0][2 * len(xs) + k].cpu().numpy(), gsum[k].cpu().numpy()) < 1e-6


def test_module_matches_torch_state_dict_and_values():
    ref = torch.nn.GroupNorm(32, 256).cuda()
    with torch.no_grad():
        ref.weight.normal_(1, 0.2)
        ref.bias.normal_(0, 0.2)
    mod = L.GroupNormReLU(32, 256).cuda()
    mod.load_state_dict(ref.state_dict())
    assert list(mod.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(2, 256, 13, 21, device="cuda")
    assert rel_err(mod(x).detach().cpu().numpy(), torch.relu(ref(x)).detach().cpu().numpy()) < 1e-5


def test_errors():
    with pytest.raises(NotImplementedError):
        L.group_norm_relu(torch.randn(1, 8, 2, 2), 2, torch.ones(8), torch.zeros(8))
    with pytest.raises(ValueError):
        L.group_norm_relu(torch.randn(1, 8, 2, 2, device="cuda"), 3, torch.ones(8, device="cuda"), torch.zeros(8, device="cuda"))
