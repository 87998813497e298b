"""The GroupNorm + ReLU oracle against the reference's implementation of those layers (torch.nn.GroupNorm + nn.ReLU,
reppointsv2.py:644-675), forward and autograd gradients, float64."""
import numpy as np
import pytest
import torch

from oracle import gn as ogn


@pytest.mark.parametrize("shape", [(2, 64, 7, 11, 8), (1, 256, 13, 21, 32), (3, 32, 5, 5, 32), (2, 16, 9, 4, 1)])
@pytest.mark.parametrize("relu", [True, False])
def test_oracle_matches_torch_group_norm_relu(shape, relu):
    N, C, H, W, G = shape
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(N, C, H, W, generator=g, dtype=torch.float64) * 1.7 + 0.3).requires_grad_()
    gamma = (torch.randn(C, generator=g, dtype=torch.float64) * 0.5 + 1.0).requires_grad_()
    beta = (torch.randn(C, generator=g, dtype=torch.float64) * 0.2).requires_grad_()
    gy = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    layer = torch.nn.GroupNorm(G, C).double()
    with torch.no_grad():
        layer.weight.copy_(gamma)
        layer.bias.copy_(beta)
    y = layer(x)
    if relu:
        y = torch.nn.ReLU(inplace=True)(y)
    y.backward(gy)
    yo, _, _ = ogn.forward(x.detach().numpy(), gamma.detach().numpy(), beta.detach().numpy(), G, layer.eps, relu)
    go = ogn.backward(x.detach().numpy(), gamma.detach().numpy(), beta.detach().numpy(), G, gy.numpy(), layer.eps, relu)
    assert np.allclose(yo, y.detach().numpy(), rtol=1e-11, atol=1e-12)
    assert np.allclose(go["grad_x"], x.grad.numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(go["grad_gamma"], layer.weight.grad.numpy(), rtol=1e-10, atol=1e-11)
    assert np.allclose(go["grad_beta"], layer.bias.grad.numpy(), rtol=1e-10, atol=1e-11)
