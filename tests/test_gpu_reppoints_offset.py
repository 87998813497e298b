"""GPU parity of the fused RepPoints DCN offset construction (C ABI) against the oracle: bit-exact forward and
backward, both channel conventions (reppointsv2.py: no flip; rpd.py: (x, y) -> (y, x))."""
import numpy as np
import pytest
import torch

from slenderobjdet_b200 import layers as L
from oracle import reppoints as orp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("shape", [(2, 18, 25, 42), (1, 18, 100, 168), (3, 50, 7, 11), (2, 2, 3, 5)])
def test_forward_backward_bit_exact(shape, flip):
    g = torch.Generator().manual_seed(4)
    pts = torch.randn(*shape, generator=g) * 4
    go = torch.randn(*shape, generator=g)
    p = pts.cuda().requires_grad_()
    out = L.reppoints_dcn_offset(p, 0.1, flip)
    out.backward(go.cuda())
    assert np.array_equal(out.detach().cpu().numpy(), orp.dcn_offset(pts.numpy(), 0.1, flip))
    assert np.array_equal(p.grad.cpu().numpy(), orp.dcn_offset_grad(go.numpy(), 0.1, flip))


def test_matches_reference_expression_on_device_and_feeds_deform_conv():
    import slenderobjdet_b200 as sdb
    g = torch.Generator().manual_seed(5)
    pts = (torch.randn(2, 18, 13, 21, generator=g) * 2).cuda()
    base = L.dcn_base_offset(9, device="cuda")
    ref = ((1 - 0.1) * pts.detach() + 0.1 * pts) - base
    off = L.reppoints_dcn_offset(pts, 0.1)
    assert torch.equal(off, ref)
    conv = sdb.DeformConv(64, 64, 3, 1, 1).cuda()
    x = torch.randn(2, 64, 13, 21, device="cuda")
    assert torch.equal(conv(x, off), conv(x, ref))


def test_errors():
    with pytest.raises(ValueError):
        L.reppoints_dcn_offset(torch.zeros(1, 6, 4, 4, device="cuda"))   # 3 points: not a square number
    with pytest.raises(NotImplementedError):
        L.reppoints_dcn_offset(torch.zeros(1, 18, 4, 4))
