"""CPU: pin the C DCN oracle against the reference's known-answer test and torchvision goldens."""
import numpy as np
import pytest

from conftest import GOLDEN, rel_err
from oracle import dcn as odcn


def test_known_answer_reference_test():
    """/root/reference/tests/test_deformable_conv.py:85-87 -- three asserts at < 1e-5."""
    z = np.load(f"{GOLDEN}/dcn_known_answer.npz")
    y1 = odcn.forward(z["x"], z["offsets_1"], z["weight"], stride=1, padding=1)
    y2 = odcn.forward(z["x"], z["offsets_2"], z["weight"], stride=1, padding=1)
    assert np.all(np.abs(y2 - z["expected_conv"]) < 1e-5)
    assert np.all(np.abs(y2 - z["expected_dconv_zero"]) < 1e-5)
    assert np.all(np.abs(y1 - z["expected_dconv_grid"]) < 1e-5)


def _cfg(c):
    sh, sw, ph, pw, dh, dw, g, dg = [int(v) for v in c["cfg"]]
    return dict(stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw), groups=g, deformable_groups=dg)


CASES = ["v1_basic", "v1_big_offsets", "v1_groups_dg", "v1_stride2_dil2", "v1_k1", "v1_k5x3", "v1_c64",
         "v2_basic", "v2_nobias_dg2", "v2_c64"]


@pytest.mark.parametrize("name", CASES)
def test_forward_vs_torchvision_golden(dcn_cases, name):
    c = dcn_cases[name]
    y = odcn.forward(c["x"], c["offset"], c["weight"], mask=c.get("mask"), bias=c.get("bias"), **_cfg(c))
    assert y.shape == c["out"].shape
    assert rel_err(y, c["out"]) < 2e-6


@pytest.mark.parametrize("name", CASES)
def test_backward_vs_torchvision_golden(dcn_cases, name):
    c = dcn_cases[name]
    g = odcn.backward(c["x"], c["offset"], c["weight"], c["grad_out"], mask=c.get("mask"),
                      with_bias="bias" in c, **_cfg(c))
    assert rel_err(g["grad_x"], c["grad_x"]) < 5e-6
    assert rel_err(g["grad_offset"], c["grad_offset"]) < 5e-6
    assert rel_err(g["grad_weight"], c["grad_weight"]) < 5e-6
    if "mask" in c:
        assert rel_err(g["grad_mask"], c["grad_mask"]) < 5e-6
    if "bias" in c:
        assert rel_err(g["grad_bias"], c["grad_bias"]) < 5e-6


def test_zero_offset_equals_conv2d():
    torch = pytest.importorskip("torch")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 9, 8, generator=g)
    w = torch.randn(4, 6, 3, 3, generator=g)
    off = torch.zeros(2, 18, 9, 8)
    y = odcn.forward(x.numpy(), off.numpy(), w.numpy(), stride=1, padding=1)
    ref = torch.nn.functional.conv2d(x, w, padding=1).numpy()
    assert rel_err(y, ref) < 2e-6


def test_all_taps_outside_gives_zero_and_zero_grads():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 4, 5, 5)).astype(np.float32)
    w = rng.standard_normal((3, 4, 3, 3)).astype(np.float32)
    off = np.full((1, 18, 5, 5), 100.0, np.float32)
    y = odcn.forward(x, off, w, stride=1, padding=1)
    assert np.all(y == 0)
    g = odcn.backward(x, off, w, np.ones_like(y), stride=1, padding=1)
    assert np.all(g["grad_x"] == 0) and np.all(g["grad_offset"] == 0) and np.all(g["grad_weight"] == 0)
