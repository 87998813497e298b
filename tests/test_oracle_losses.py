"""CPU: pin the loss oracle against the reference's own Python (goldens) and torchvision."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import losses as ol


def test_focal_vs_torchvision_golden(loss_cases):
    c = loss_cases["focal"]
    s, g = ol.sigmoid_focal_loss(c["logits"], c["cls"], alpha=0.25, gamma=2.0)
    assert abs(float(s) - float(c["loss"])) / float(c["loss"]) < 1e-5
    assert rel_err(g.numpy(), c["grad"]) < 1e-5


@pytest.mark.parametrize("form", ["ltrb", "xyxy"])
@pytest.mark.parametrize("lt", ["iou", "linear_iou", "giou"])
@pytest.mark.parametrize("use_w", [0, 1])
def test_iou_loss_vs_reference_golden(loss_cases, form, lt, use_w):
    d, c = loss_cases[form], loss_cases[f"{form}_{lt}_{use_w}"]
    s, g = ol.iou_loss(d["pred"], d["target"], d["weight"] if use_w else None, loss_type=lt, form=form)
    assert abs(float(s) - float(c["loss"])) / abs(float(c["loss"])) < 2e-5
    assert rel_err(g.numpy(), c["grad"]) < 2e-4  # reference grad is float32 autograd


@pytest.mark.parametrize("beta", [0.11, 0.0])
@pytest.mark.parametrize("use_w", [0, 1])
def test_smooth_l1_vs_reference_golden(loss_cases, beta, use_w):
    d, c = loss_cases["sl1"], loss_cases[f"sl1_{beta}_{use_w}"]
    s, g = ol.smooth_l1_loss(d["pred"], d["target"], beta, d["weight"] if use_w else None)
    assert abs(float(s) - float(c["loss"])) / abs(float(c["loss"])) < 1e-5
    assert rel_err(g.numpy(), c["grad"]) < 1e-5


def test_giou_vs_torchvision_golden(loss_cases):
    c = loss_cases["giou"]
    s, g = ol.giou_loss(c["pred"], c["target"])
    assert abs(float(s) - float(c["loss"])) / abs(float(c["loss"])) < 1e-5
    assert rel_err(g.numpy(), c["grad"]) < 1e-4


def test_fcos_rpd_losses_oracle_vs_reference(target_cases):
    """FCOSRepPoints.losses EXECUTED from the reference (fcos_rpd_s1_topk.py:249-317, gen_target_golden.py) on the
    targets its own get_ground_truth produced: four loss values and the gradients of their sum with respect to the
    four prediction tensors.  (The reference ran in float32; the oracle is float64.)"""
    t, c = target_cases["rpd_small"], target_cases["rpd_small_loss"]
    losses, grads = ol.fcos_rpd_losses(t["init_classes"], t["init_reg"], t["refine_classes"], t["refine_reg"], c["logits"],
                                       c["box_init"], c["box_ref"], c["ctr"], c["strides"], t["topk"], 80)
    for k in ("cls_loss", "reg_loss_init", "reg_loss", "centerness_loss"):
        assert abs(float(losses[k]) - float(c[k])) / abs(float(c[k])) < 2e-5, k
    assert rel_err(grads["pred_class_logits"].numpy(), c["g_logits"]) < 2e-5
    assert rel_err(grads["pred_box_reg_init"].numpy(), c["g_box_init"]) < 2e-4
    assert rel_err(grads["pred_box_reg"].numpy(), c["g_box_ref"]) < 2e-5
    assert rel_err(grads["pred_center_score"].numpy(), c["g_ctr"]) < 2e-5
