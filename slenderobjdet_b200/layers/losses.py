"""Fused dense-head losses behind the reference's call signatures.

  sigmoid_focal_loss(_jit)   fvcore.nn signature used at reppointsv2.py:307-312, fcos.py:293-297
  iou_loss / box_iou_loss    /root/reference/slender_det/layers/iou_loss.py:4-77
  smooth_l1_loss(_with_weight)  fvcore signature / sd/layers/smooth_l1_loss_with_weight.py:3-17
  giou_loss                  fvcore signature used at meta/heads/anchor_head.py:369-376

Each is ONE kernel launch computing the sum-reduced loss and its gradient in the same pass (the
reference runs ~5-30 eager kernels and, for focal, materialises a dense one-hot target).  The
fused kernels implement ``reduction="sum"`` (what every call site in the reference uses) on CUDA
float32 tensors; anything else raises -- there is no CPU / eager fallback.
"""
import torch
from torch.autograd import Function

from .. import _lib


def _req(cond, msg):
    if not cond:
        raise RuntimeError("slender_b200 losses: " + msg)


def _f32c(t):
    return t.detach().float().contiguous()


class _Focal(Function):
    @staticmethod
    def forward(ctx, logits, class_idx, alpha, gamma):
        _req(logits.is_cuda, "CUDA tensors only (no CPU fallback)")
        _req(logits.dim() == 2, "logits must be [R, K]")
        x = _f32c(logits)
        idx = class_idx.detach().to(torch.int64).contiguous()
        _req(idx.shape == (x.shape[0],), "class_idx must be [R]")
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if logits.requires_grad else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sdb_sigmoid_focal_loss(_lib.ptr(x), _lib.ptr(idx), x.shape[0], x.shape[1],
                                                         float(alpha), float(gamma), 1.0, _lib.ptr(loss),
                                                         _lib.ptr(grad), _lib.stream_ptr(x.device)))
        ctx.grad = grad
        ctx.dtype = logits.dtype
        return loss.to(logits.dtype)

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g).to(ctx.dtype), None, None, None


def sigmoid_focal_loss_from_class_idx(logits, class_idx, alpha=-1.0, gamma=2.0):
    """sum-reduced focal loss; ``class_idx[r]`` in [0,K) = foreground class of row r, else background."""
    return _Focal.apply(logits, class_idx, alpha, gamma)


def sigmoid_focal_loss(inputs, targets, alpha=-1, gamma=2, reduction="none"):
    """fvcore signature.  ``targets`` is the one-hot tensor the reference heads build; rows must hold
    at most one 1 (true for every call site).  Only reduction='sum' is fused."""
    _req(reduction == "sum", "only reduction='sum' is implemented (the reference's call sites use it)")
    _req(inputs.shape == targets.shape, "inputs / targets shape mismatch")
    K = inputs.shape[-1]
    t = targets.reshape(-1, K)
    val, idx = t.max(dim=1)
    idx = torch.where(val > 0, idx, torch.full_like(idx, K))
    return _Focal.apply(inputs.reshape(-1, K), idx, alpha, gamma)


sigmoid_focal_loss_jit = sigmoid_focal_loss


class _BoxLoss(Function):
    @staticmethod
    def forward(ctx, pred, target, weight, kind, form, beta):
        _req(pred.is_cuda, "CUDA tensors only (no CPU fallback)")
        _req(pred.dim() == 2 and pred.shape[1] == 4 and pred.shape == target.shape, "pred / target must be [R, 4]")
        p, t = _f32c(pred), _f32c(target)
        w = None if weight is None else _f32c(weight)
        if w is not None:
            _req(w.shape == (p.shape[0],), "weight must be [R]")
        loss = torch.zeros((), dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p) if pred.requires_grad else None
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().sdb_box_reg_loss(_lib.ptr(p), _lib.ptr(t), _lib.ptr(w), p.shape[0], kind, form,
                                                   float(beta), 1.0, _lib.ptr(loss), _lib.ptr(grad),
                                                   _lib.stream_ptr(p.device)))
        ctx.grad = grad
        ctx.dtype = pred.dtype
        return loss.to(pred.dtype)

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g).to(ctx.dtype), None, None, None, None, None


_KINDS = {"iou": _lib.SDB_LOSS_IOU, "linear_iou": _lib.SDB_LOSS_LINEAR_IOU, "giou": _lib.SDB_LOSS_GIOU}


def iou_loss(pred, target, weight=None, loss_type="iou"):
    """(l, t, r, b) distances; sd/layers/iou_loss.py:4-37."""
    if loss_type not in _KINDS:
        raise NotImplementedError
    if weight is None:
        assert pred.numel() != 0
    return _BoxLoss.apply(pred, target, weight, _KINDS[loss_type], _lib.SDB_BOX_LTRB, 0.0)


def box_iou_loss(pred, target, weight=None, loss_type="iou"):
    """(x1, y1, x2, y2) boxes; sd/layers/iou_loss.py:40-77."""
    if loss_type not in _KINDS:
        raise NotImplementedError
    if weight is None:
        assert pred.numel() != 0
    return _BoxLoss.apply(pred, target, weight, _KINDS[loss_type], _lib.SDB_BOX_XYXY, 0.0)


def smooth_l1_loss_with_weight(input, target, weight, beta, reduction="none"):
    """sd/layers/smooth_l1_loss_with_weight.py:3-17 for [R,4] tensors, reduction='sum'."""
    _req(reduction == "sum", "only reduction='sum' is implemented")
    return _BoxLoss.apply(input, target, weight, _lib.SDB_LOSS_SMOOTH_L1, _lib.SDB_BOX_XYXY, beta)


def smooth_l1_loss(input, target, beta, reduction="none"):
    """fvcore signature (reppointsv2.py:314-320), [R,4] tensors, reduction='sum'."""
    return smooth_l1_loss_with_weight(input, target, None, beta, reduction)


def giou_loss(boxes1, boxes2, reduction="none", eps=1e-7):
    """fvcore signature (anchor_head.py:369-376), xyxy boxes, reduction='sum', eps=1e-7."""
    _req(reduction == "sum", "only reduction='sum' is implemented")
    _req(abs(eps - 1e-7) < 1e-12, "eps is fixed at 1e-7")
    return _BoxLoss.apply(boxes1, boxes2, None, _lib.SDB_LOSS_GIOU_FVCORE, _lib.SDB_BOX_XYXY, 0.0)


def _centerness(reg_targets, slender):
    _req(reg_targets.dim() == 2 and reg_targets.shape[1] == 4, "reg_targets must be [R, 4]")
    if not reg_targets.is_cuda:
        raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
    r = _f32c(reg_targets.detach())
    out = torch.empty((r.shape[0],), dtype=torch.float32, device=r.device)
    fn = _lib.lib().sdb_slender_centerness_targets if slender else _lib.lib().sdb_centerness_targets
    with torch.cuda.device(r.device):
        _lib.check(fn(_lib.ptr(r), r.shape[0], _lib.ptr(out), _lib.stream_ptr(r.device)))
    return out.to(reg_targets.dtype)


def compute_centerness_targets(reg_targets):
    """``slender_det.modeling.meta_arch.fcos.utils.compute_centerness_targets`` (fcos/utils.py:295-300):
    ``sqrt(min(l, r) / max(l, r) * min(t, b) / max(t, b))`` for ``reg_targets [R, 4]`` in (l, t, r, b) order.
    No gradient (the reference uses it as a target)."""
    return _centerness(reg_targets, False)


def compute_slender_centerness_targets(reg_targets):
    """The ``compute_centerness_targets`` that the active FCOSRepPoints model defines for itself
    (fcos/fcos_rpd_s1_topk.py:25-55) and uses in its losses (:288, :291) and top-5 selection (:117):
    ``pow(c, min(w/h, h/w))`` with ``c`` the product above, ``w = l + r``, ``h = t + b``."""
    return _centerness(reg_targets, True)
