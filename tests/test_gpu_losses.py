"""GPU parity: fused losses vs goldens from the reference's Python / torchvision and the float64 oracle.
Floating-point kernels: rel <= 1e-4 (fp32 path tolerance of BASELINE.json)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from slenderobjdet_b200 import layers as L
from oracle import losses as ol

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _d(a):
    return torch.as_tensor(a).cuda()


def test_focal_vs_golden_and_onehot_signature(loss_cases):
    c = loss_cases["focal"]
    x = _d(c["logits"]).requires_grad_()
    loss = L.sigmoid_focal_loss_from_class_idx(x, _d(c["cls"]), alpha=0.25, gamma=2.0)
    (loss * 0.5).backward()
    assert abs(float(loss) - float(c["loss"])) / float(c["loss"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), 0.5 * c["grad"]) < TOL
    K = x.shape[1]
    t = torch.zeros_like(x)
    idx = _d(c["cls"])
    fg = idx < K
    t[fg.nonzero(as_tuple=True)[0], idx[fg]] = 1
    x2 = _d(c["logits"]).requires_grad_()
    l2 = L.sigmoid_focal_loss_jit(x2, t, alpha=0.25, gamma=2.0, reduction="sum")  # reference call signature
    assert abs(float(l2) - float(c["loss"])) / float(c["loss"]) < TOL


@pytest.mark.parametrize("gamma,alpha", [(2.0, 0.25), (1.5, -1.0), (0.0, 0.5)])
def test_focal_vs_oracle_full_size(gamma, alpha):
    """BASELINE config 4 shape: 2 images x 22 400 points x 80 classes."""
    g = torch.Generator().manual_seed(1)
    R, K = 44800, 80
    x = torch.randn(R, K, generator=g) * 2 - 4.6
    cls = torch.full((R,), K, dtype=torch.int64)
    pos = torch.randperm(R, generator=g)[:450]
    cls[pos] = torch.randint(0, K, (450,), generator=g)
    so, go = ol.sigmoid_focal_loss(x, cls, alpha, gamma)
    xd = x.cuda().requires_grad_()
    loss = L.sigmoid_focal_loss_from_class_idx(xd, cls.cuda(), alpha, gamma)
    loss.backward()
    assert abs(float(loss) - float(so)) / float(so) < TOL
    assert rel_err(xd.grad.cpu().numpy(), go.numpy()) < TOL


@pytest.mark.parametrize("form", ["ltrb", "xyxy"])
@pytest.mark.parametrize("lt", ["iou", "linear_iou", "giou"])
@pytest.mark.parametrize("use_w", [0, 1])
def test_iou_losses_vs_golden(loss_cases, form, lt, use_w):
    d, c = loss_cases[form], loss_cases[f"{form}_{lt}_{use_w}"]
    fn = L.iou_loss if form == "ltrb" else L.box_iou_loss
    p = _d(d["pred"]).requires_grad_()
    loss = fn(p, _d(d["target"]), _d(d["weight"]) if use_w else None, loss_type=lt)
    loss.backward()
    so, go = ol.iou_loss(d["pred"], d["target"], d["weight"] if use_w else None, lt, form)
    assert abs(float(loss) - float(c["loss"])) / abs(float(c["loss"])) < TOL
    assert rel_err(p.grad.cpu().numpy(), go.numpy()) < TOL


@pytest.mark.parametrize("beta", [0.11, 0.0])
@pytest.mark.parametrize("use_w", [0, 1])
def test_smooth_l1_vs_golden(loss_cases, beta, use_w):
    d, c = loss_cases["sl1"], loss_cases[f"sl1_{beta}_{use_w}"]
    p = _d(d["pred"]).requires_grad_()
    loss = L.smooth_l1_loss_with_weight(p, _d(d["target"]), _d(d["weight"]) if use_w else None, beta, reduction="sum")
    loss.backward()
    assert abs(float(loss) - float(c["loss"])) / abs(float(c["loss"])) < TOL
    assert rel_err(p.grad.cpu().numpy(), c["grad"]) < TOL


def test_fvcore_giou_vs_golden(loss_cases):
    c = loss_cases["giou"]
    p = _d(c["pred"]).requires_grad_()
    loss = L.giou_loss(p, _d(c["target"]), reduction="sum")
    loss.backward()
    assert abs(float(loss) - float(c["loss"])) / abs(float(c["loss"])) < TOL
    so, go = ol.giou_loss(c["pred"], c["target"])
    assert rel_err(p.grad.cpu().numpy(), go.numpy()) < TOL


def test_centerness_targets_bit_exact():
    """compute_centerness_targets (fcos/utils.py:295-300): bit-exact against the reference's lines on CUDA tensors,
    within float32 rounding of the float64 oracle; degenerate rows (a zero side -> 0, equal sides -> 1)."""
    g = torch.Generator().manual_seed(3)
    r = torch.rand(5000, 4, generator=g) * 200 + 1e-3
    r[0] = torch.tensor([5.0, 5.0, 5.0, 5.0])
    r[1] = torch.tensor([0.0, 3.0, 7.0, 9.0])
    out = L.compute_centerness_targets(r.cuda())
    assert out.dtype == torch.float32 and out.shape == (5000,)
    # the reference's own lines, executed where the reference executes them (float32 CUDA tensors): bit-exact.
    # (torch's vectorised CPU kernels differ from the CUDA ones in the last bit on ~1 % of rows.)
    rd = r.cuda()
    lr, tb = rd[:, [0, 2]], rd[:, [1, 3]]
    ref32 = torch.sqrt((lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0]))
    assert np.array_equal(out.cpu().numpy(), ref32.cpu().numpy())
    # and the float64 oracle restatement within float32 rounding
    ref = np.asarray(ol.centerness_targets(r), dtype=np.float64)
    assert np.max(np.abs(out.cpu().numpy().astype(np.float64) - ref)) <= 1e-6
    assert float(out[0]) == 1.0 and float(out[1]) == 0.0
    assert L.compute_centerness_targets(torch.zeros(0, 4, device="cuda")).shape == (0,)


def test_fcos_rpd_losses_match_oracle():
    """FCOSRepPoints.losses (fcos_rpd_s1_topk.py:249-317) composed from the fused kernels without host sync:
    the four loss values and the gradients of their sum against the float64 restatement with the reference's
    boolean-mask selections (rel <= 1e-4)."""
    from slenderobjdet_b200.fcos_rpd_losses import fcos_rpd_losses
    g = torch.Generator().manual_seed(17)
    N, X, K = 2, 3000, 80
    R = N * X
    icls = torch.full((R,), K, dtype=torch.long)
    fg = torch.randperm(R, generator=g)[:400]
    icls[fg] = torch.randint(0, K, (400,), generator=g)
    ireg = torch.randn(R, 4, generator=g) * 30                      # background rows: arbitrary, also negative
    ireg[fg] = torch.rand(400, 4, generator=g) * 60 + 1
    topk = torch.zeros(R, dtype=torch.bool); topk[fg[:150]] = True
    rcls = torch.full((R,), K, dtype=torch.long)
    rfg = torch.randperm(R, generator=g)[:500]
    rcls[rfg] = torch.randint(0, K, (500,), generator=g)
    rcls[torch.randperm(R, generator=g)[:100]] = -1                # centres outside the image
    rreg = torch.randn(R, 4, generator=g) * 40
    logits = torch.randn(R, K, generator=g) * 2 - 4.6
    pbi = torch.rand(R, 4, generator=g) * 60 + 0.5
    pb = torch.randn(R, 4, generator=g) * 40
    pc = torch.randn(R, generator=g)
    strides = torch.tensor([8.0, 16.0, 32.0])[torch.randint(0, 3, (R,), generator=g)]
    d = lambda t: t.cuda()
    preds = [d(logits).requires_grad_(), d(pbi).requires_grad_(), d(pb).requires_grad_(), d(pc).requires_grad_()]
    out = fcos_rpd_losses(d(icls).view(N, X), d(ireg).view(N, X, 4), d(rcls).view(N, X), d(rreg).view(N, X, 4), *preds,
                          d(strides), d(topk).view(N, X), K)
    sum(out.values()).backward()
    lo, go = ol.fcos_rpd_losses(icls, ireg, rcls, rreg, logits, pbi, pb, pc, strides, topk, K)
    for k in ("cls_loss", "reg_loss_init", "reg_loss", "centerness_loss"):
        assert abs(float(out[k]) - float(lo[k])) <= 1e-4 * abs(float(lo[k])), (k, float(out[k]), float(lo[k]))
    for p, k in zip(preds, ("pred_class_logits", "pred_box_reg_init", "pred_box_reg", "pred_center_score")):
        assert rel_err(p.grad.cpu().numpy(), go[k].numpy()) < 1e-4, k


def test_fcos_rpd_losses_vs_reference(target_cases):
    """The same composition against FCOSRepPoints.losses EXECUTED from the reference on the targets its own
    get_ground_truth produced (tests/golden/gen_target_golden.py; fvcore's focal / smooth-L1 stood in by the
    published formulas): the four losses and the gradients of their sum.  The reference ran in float32 on the CPU."""
    from slenderobjdet_b200.fcos_rpd_losses import fcos_rpd_losses
    t, c = target_cases["rpd_small"], target_cases["rpd_small_loss"]
    d = lambda a: torch.as_tensor(a).cuda()
    preds = [d(c["logits"]).requires_grad_(), d(c["box_init"]).requires_grad_(), d(c["box_ref"]).requires_grad_(),
             d(c["ctr"]).requires_grad_()]
    out = fcos_rpd_losses(d(t["init_classes"]), d(t["init_reg"]), d(t["refine_classes"]), d(t["refine_reg"]), *preds,
                          d(c["strides"]), d(t["topk"]), 80)
    sum(out.values()).backward()
    for k in ("cls_loss", "reg_loss_init", "reg_loss", "centerness_loss"):
        assert abs(float(out[k]) - float(c[k])) <= 5e-5 * abs(float(c[k])), (k, float(out[k]), float(c[k]))
    for p, k, tol in zip(preds, ("g_logits", "g_box_init", "g_box_ref", "g_ctr"), (1e-4, 3e-4, 1e-4, 1e-4)):
        assert rel_err(p.grad.cpu().numpy(), c[k]) < tol, k


def test_slender_centerness_vs_reference(target_cases):
    """compute_centerness_targets as the FCOSRepPoints module defines it for itself (fcos_rpd_s1_topk.py:25-55):
    pow(c, min(w/h, h/w)), against the reference's output (CUDA powf vs torch CPU pow: a few ulp)."""
    c = target_cases["ctr"]
    out = L.compute_slender_centerness_targets(torch.as_tensor(c["ltrb"]).cuda()).cpu().numpy()
    assert np.allclose(out, c["slender"], rtol=2e-6, atol=0)
    out2 = L.compute_centerness_targets(torch.as_tensor(c["ltrb"]).cuda()).cpu().numpy()
    assert np.allclose(out2, c["fcos"], rtol=2e-7, atol=0)
