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
