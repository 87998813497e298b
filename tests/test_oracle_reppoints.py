"""Oracle for the RepPoints DCN offset construction against the reference's own lines run with torch on CPU."""
import numpy as np
import torch

from oracle import reppoints as orp


def _reference_lines(pts, gradient_mul, flip):
    """reppointsv2.py:638-642, 742-744 / rpd.py:105-110, 624-635 verbatim in torch."""
    num_points = pts.shape[1] // 2
    dcn_kernel = int(np.sqrt(num_points))
    dcn_pad = int((dcn_kernel - 1) / 2)
    dcn_base = np.arange(-dcn_pad, dcn_pad + 1).astype(np.float64)
    dcn_base_y = np.repeat(dcn_base, dcn_kernel)
    dcn_base_x = np.tile(dcn_base, dcn_kernel)
    dcn_base_offset = torch.tensor(np.stack([dcn_base_y, dcn_base_x], axis=1).reshape((-1))).view(1, -1, 1, 1)
    p = pts.clone().requires_grad_()
    gm = (1 - gradient_mul) * p.detach() + gradient_mul * p
    if flip:
        gm = gm.reshape(gm.size(0), num_points, 2, *gm.shape[-2:]).flip(2).reshape(-1, 2 * num_points, *gm.shape[-2:])
    out = gm - dcn_base_offset.type_as(p)
    return p, out


def test_oracle_matches_reference_lines():
    g = torch.Generator().manual_seed(0)
    for k, flip in ((9, False), (9, True), (25, False), (1, True)):
        pts = torch.randn(2, 2 * k, 5, 7, generator=g) * 3
        p, out = _reference_lines(pts, 0.1, flip)
        assert np.array_equal(orp.dcn_offset(pts.numpy(), 0.1, flip), out.detach().numpy())
        go = torch.randn(out.shape, generator=g)
        out.backward(go)
        assert np.array_equal(orp.dcn_offset_grad(go.numpy(), 0.1, flip), p.grad.numpy())


def test_base_offset_layout():
    assert orp.dcn_base_offset(9).tolist() == [-1, -1, -1, 0, -1, 1, 0, -1, 0, 0, 0, 1, 1, -1, 1, 0, 1, 1]
