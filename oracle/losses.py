"""CPU restatement of the reference's dense-head losses -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Formulas followed:
  iou_loss / box_iou_loss ........ /root/reference/slender_det/layers/iou_loss.py:4-37, :40-77
  smooth_l1_loss_with_weight ..... /root/reference/slender_det/layers/smooth_l1_loss_with_weight.py:3-17
  sigmoid_focal_loss_jit, smooth_l1_loss, giou_loss: third-party ``fvcore`` (pinned only as
    ``fvcore>=0.1.1``: /root/reference/setup.py:106, detectron2/setup.py:194), source NOT under
    /root/reference and not installed.  Restated from the published fvcore.nn formulas
    (focal_loss.py, smooth_l1_loss.py, giou_loss.py); call sites: reppointsv2.py:307-320,
    fcos.py:293-297, anchor_head.py:369-376.  PARITY UNPINNED by the reference's own tests; pinned
    here against torchvision.ops.sigmoid_focal_loss / generalized_box_iou_loss
    (tests/test_oracle_losses.py), which implement the same published formulas.

Everything is evaluated in float64 torch on CPU; gradients come from autograd on the restated
forward, so the CUDA kernels' hand-derived gradients are checked against an independent path.
Each function returns (loss_sum: float64 scalar tensor, grad wrt the first argument).
"""
import torch


def _prep(*ts):
    return [None if t is None else torch.as_tensor(t).detach().to(torch.float64).cpu() for t in ts]


def sigmoid_focal_loss(logits, class_idx, alpha=0.25, gamma=2.0):
    """logits [R,K]; class_idx [R] int (k in [0,K) = foreground class, anything else = background).

    Equals fvcore sigmoid_focal_loss_jit(logits, one_hot(class_idx), alpha, gamma, "sum") -- the
    one-hot target built at reppointsv2.py:294-295 / fcos.py:289-292.
    """
    (x,) = _prep(logits)
    x.requires_grad_(True)
    R, K = x.shape
    idx = torch.as_tensor(class_idx).long().cpu()
    t = torch.zeros_like(x)
    fg = (idx >= 0) & (idx < K)
    t[fg.nonzero(as_tuple=True)[0], idx[fg]] = 1.0
    p = torch.sigmoid(x)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(x, t, reduction="none")
    p_t = p * t + (1 - p) * (1 - t)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * t + (1 - alpha) * (1 - t)) * loss
    s = loss.sum()
    (g,) = torch.autograd.grad(s, x)
    return s.detach(), g


def smooth_l1_loss(pred, target, beta, weight=None):
    """sum-reduced smooth-L1; ``weight`` [R] multiplies each row (smooth_l1_loss_with_weight)."""
    x, y, w = _prep(pred, target, weight)
    x.requires_grad_(True)
    n = (x - y).abs()
    loss = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if w is not None:
        loss = loss * w[:, None]
    s = loss.sum()
    (g,) = torch.autograd.grad(s, x)
    return s.detach(), g


def iou_loss(pred, target, weight=None, loss_type="iou", form="ltrb"):
    """iou_loss (form='ltrb') / box_iou_loss (form='xyxy'), loss_type in {iou, linear_iou, giou}."""
    p, t, w = _prep(pred, target, weight)
    p.requires_grad_(True)
    pl, pt, pr, pb = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    tl, tt, tr, tb = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    if form == "ltrb":
        t_area = (tl + tr) * (tt + tb)
        p_area = (pl + pr) * (pt + pb)
        w_i = torch.min(pl, tl) + torch.min(pr, tr)
        g_w = torch.max(pl, tl) + torch.max(pr, tr)
        h_i = torch.min(pb, tb) + torch.min(pt, tt)
        g_h = torch.max(pb, tb) + torch.max(pt, tt)
    elif form == "xyxy":
        t_area = (tr - tl) * (tb - tt)
        p_area = (pr - pl) * (pb - pt)
        w_i = torch.min(pr, tr) - torch.max(pl, tl)
        g_w = torch.max(pr, tr) - torch.min(pl, tl)
        h_i = torch.min(pb, tb) - torch.max(pt, tt)
        g_h = torch.max(pb, tb) - torch.min(pt, tt)
    else:
        raise ValueError(form)
    ac = g_w * g_h + 1e-7
    a_i = w_i * h_i
    a_u = t_area + p_area - a_i
    ious = (a_i + 1.0) / (a_u + 1.0)
    gious = ious - (ac - a_u) / ac
    if loss_type == "iou":
        losses = -torch.log(ious)
    elif loss_type == "linear_iou":
        losses = 1 - ious
    elif loss_type == "giou":
        losses = 1 - gious
    else:
        raise NotImplementedError(loss_type)
    s = (losses * w).sum() if w is not None else losses.sum()
    (g,) = torch.autograd.grad(s, p)
    return s.detach(), g


def giou_loss(pred, target, eps=1e-7):
    """fvcore giou_loss(boxes1, boxes2, reduction='sum'), xyxy boxes."""
    p, t = _prep(pred, target)
    p.requires_grad_(True)
    x1, y1, x2, y2 = p.unbind(-1)
    x1g, y1g, x2g, y2g = t.unbind(-1)
    xkis1, ykis1 = torch.max(x1, x1g), torch.max(y1, y1g)
    xkis2, ykis2 = torch.min(x2, x2g), torch.min(y2, y2g)
    intsct = torch.zeros_like(x1)
    m = (ykis2 > ykis1) & (xkis2 > xkis1)
    intsct = torch.where(m, (xkis2 - xkis1) * (ykis2 - ykis1), intsct)
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - intsct
    iou = intsct / (union + eps)
    xc1, yc1 = torch.min(x1, x1g), torch.min(y1, y1g)
    xc2, yc2 = torch.max(x2, x2g), torch.max(y2, y2g)
    area_c = (xc2 - xc1) * (yc2 - yc1)
    miou = iou - ((area_c - union) / (area_c + eps))
    s = (1 - miou).sum()
    (g,) = torch.autograd.grad(s, p)
    return s.detach(), g


def centerness_targets(ltrb):
    """compute_centerness_targets: /root/reference/slender_det/modeling/meta_arch/fcos/utils.py:295-300."""
    (r,) = _prep(ltrb)
    lr = r[:, [0, 2]]
    tb = r[:, [1, 3]]
    c = (lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0])
    return torch.sqrt(c)


def slender_centerness_targets(ltrb):
    """The FCOSRepPoints module's own compute_centerness_targets
    (/root/reference/slender_det/modeling/meta_arch/fcos/fcos_rpd_s1_topk.py:25-55): pow(c, gt_ratio) with
    c as above and gt_ratio = min((l+r)/(t+b), (t+b)/(l+r)) -- the slender-object exponent.  It shadows the
    fcos/utils.py function inside that module (:288, :291 and the top-5 loop :117)."""
    (r,) = _prep(ltrb)
    lr = r[:, [0, 2]]
    tb = r[:, [1, 3]]
    ratio1 = (r[:, 0] + r[:, 2]) / (r[:, 1] + r[:, 3])
    ratio = torch.stack((ratio1, 1 / ratio1), dim=1).min(dim=1)[0]
    c = (lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0])
    return torch.pow(c, ratio)


def fcos_rpd_losses(init_gt_classes, init_reg_targets, refine_gt_classes, refine_reg_targets, pred_class_logits,
                    pred_box_reg_init, pred_box_reg, pred_center_score, strides, topk_locations, num_classes,
                    alpha=0.25, gamma=2.0, iou_loss_type="iou"):
    """FCOSRepPoints.losses restated for one process (fcos_rpd_s1_topk.py:249-317), float64 on the CPU, with the
    reference's boolean-mask selections and normalisers.  -> (dict of the four losses, dict of the gradients of
    their SUM with respect to the four prediction tensors)."""
    K = num_classes
    icls = torch.as_tensor(init_gt_classes).flatten().long().cpu()
    rcls = torch.as_tensor(refine_gt_classes).flatten().long().cpu()
    (ireg, rreg, logits, pbi, pb, pc, st) = _prep(torch.as_tensor(init_reg_targets).reshape(-1, 4),
                                                  torch.as_tensor(refine_reg_targets).reshape(-1, 4), pred_class_logits,
                                                  pred_box_reg_init, pred_box_reg, pred_center_score, strides)
    topk = torch.as_tensor(topk_locations).reshape(-1).bool().cpu()
    ifg = (icls >= 0) & (icls != K)                                                 # :261
    rfg = (rcls >= 0) & (rcls != K)                                                 # :272
    init_num = max(float(ifg.sum()), 1.0)                                           # :266-267 (one process)
    ref_num = max(float(rfg.sum()), 1.0)                                            # :277-278
    cls_idx = torch.where(rfg, rcls, torch.full_like(rcls, K))
    cls_sum, g_logits = sigmoid_focal_loss(logits, cls_idx, alpha, gamma)           # :283-287
    gt_center = slender_centerness_targets(ireg[ifg])                               # :288 (the module's own pow form, :25-55)
    topk_center = slender_centerness_targets(ireg[topk])                            # :291
    sum_topk = float(topk_center.sum())                                             # :293-294
    ri_sum, g_pbi_sel = iou_loss(pbi[topk], ireg[topk], topk_center, iou_loss_type, "ltrb")    # :298-301
    norm = (st[rfg] * 4).unsqueeze(-1)                                              # :303
    rg_sum, g_pb_sel = smooth_l1_loss(pb[rfg] / norm, rreg[rfg] / norm, 0.11)       # :304-307
    x = pc[ifg].clone().requires_grad_(True)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(x, gt_center, reduction="sum")     # :313-315
    (g_pc_sel,) = torch.autograd.grad(ce, x)
    losses = dict(cls_loss=cls_sum / ref_num, reg_loss_init=ri_sum / sum_topk, reg_loss=rg_sum / max(1, ref_num),
                  centerness_loss=ce.detach() / init_num)
    g_pbi = torch.zeros_like(pbi); g_pbi[topk] = g_pbi_sel / sum_topk
    g_pb = torch.zeros_like(pb); g_pb[rfg] = g_pb_sel / norm / max(1, ref_num)
    g_pc = torch.zeros_like(pc); g_pc[ifg] = g_pc_sel / init_num
    grads = dict(pred_class_logits=g_logits / ref_num, pred_box_reg_init=g_pbi, pred_box_reg=g_pb, pred_center_score=g_pc)
    return losses, grads
