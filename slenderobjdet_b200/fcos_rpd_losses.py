"""``FCOSRepPoints.losses`` (/root/reference/slender_det/modeling/meta_arch/fcos/fcos_rpd_s1_topk.py:249-317) on the
fused loss kernels, without host synchronisation.

The reference selects rows with boolean masks / ``nonzero`` and reads five normalisers back with
``reduce_sum(...).item()``.  Here every selection is a weight (rows that are not selected
get weight 0 and benign inputs, so they contribute neither value nor gradient), and the normalisers stay on
the device (all-reduced across ranks when ``torch.distributed`` is initialised, exactly where the reference
calls ``reduce_sum``)."""
import torch
import torch.nn.functional as F

from .layers import losses as LL


def _reduce_sum(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def _num_gpus():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def fcos_rpd_losses(init_gt_classes, init_reg_targets, refine_gt_classes, refine_reg_targets, pred_class_logits,
                    pred_box_reg_init, pred_box_reg, pred_center_score, strides, topk_locations, num_classes,
                    focal_loss_alpha=0.25, focal_loss_gamma=2.0, iou_loss_type="iou"):
    """Predictions already permuted and concatenated to [N*X, K], [N*X, 4], [N*X, 4], [N*X]
    (``permute_and_concat``, :255-256); ``strides`` [N*X] (already repeated per image, :253).
    -> dict(cls_loss, reg_loss_init, reg_loss, centerness_loss), the reference's four terms."""
    K = num_classes
    ng = float(_num_gpus())
    init_cls = init_gt_classes.flatten()
    init_reg = init_reg_targets.reshape(-1, 4)
    ref_cls = refine_gt_classes.flatten()
    ref_reg = refine_reg_targets.reshape(-1, 4)
    topk = topk_locations.reshape(-1)
    init_fg = (init_cls >= 0) & (init_cls != K)                                     # :261
    ref_fg = (ref_cls >= 0) & (ref_cls != K)                                        # :272
    init_num_pos = torch.clamp(_reduce_sum(init_fg.sum().float().reshape(1)) / ng, min=1.0)      # :266-267
    ref_num_pos = torch.clamp(_reduce_sum(ref_fg.sum().float().reshape(1)) / ng, min=1.0)        # :277-278

    # classification: one-hot target rows only where refine is foreground (:280-287); -1 / K rows are all-zero rows
    cls_idx = torch.where(ref_fg, ref_cls, torch.full_like(ref_cls, K))
    cls_loss = LL.sigmoid_focal_loss_from_class_idx(pred_class_logits, cls_idx, focal_loss_alpha, focal_loss_gamma) / ref_num_pos

    # centerness targets of the stage-1 foreground / top-k rows (:288-299), as per-row weights.  Inside this model
    # `compute_centerness_targets` is the module's OWN pow(c, min(w/h, h/w)) (:25-55), not fcos/utils.py's sqrt(c)
    ctr = LL.compute_slender_centerness_targets(torch.where(init_fg[:, None] | topk[:, None], init_reg, torch.ones_like(init_reg)))
    w_fg = torch.where(init_fg, ctr, torch.zeros_like(ctr))
    w_topk = torch.where(topk, ctr, torch.zeros_like(ctr))
    sum_topk = _reduce_sum(w_topk.sum().reshape(1)) / ng                            # :293-294
    one4 = torch.ones_like(init_reg)
    reg_loss_init = LL.iou_loss(torch.where(topk[:, None], pred_box_reg_init, one4),
                                torch.where(topk[:, None], init_reg, one4), w_topk, loss_type=iou_loss_type) / sum_topk

    # refine regression: smooth-L1 on stride-normalised coordinates of the refine foreground rows (:303-307)
    norm = (strides * 4).unsqueeze(-1)
    w_ref = ref_fg.to(pred_box_reg.dtype)
    zero4 = torch.zeros_like(pred_box_reg)
    reg_loss = LL.smooth_l1_loss_with_weight(torch.where(ref_fg[:, None], pred_box_reg / norm, zero4),
                                             torch.where(ref_fg[:, None], ref_reg / norm, zero4), w_ref, 0.11,
                                             reduction="sum") / torch.clamp(ref_num_pos, min=1.0)

    # centerness: BCE with logits on the stage-1 foreground rows (:313-315)
    centerness_loss = F.binary_cross_entropy_with_logits(pred_center_score, w_fg, weight=init_fg.to(pred_center_score.dtype),
                                                         reduction="sum") / init_num_pos
    return dict(cls_loss=cls_loss.reshape(()), reg_loss_init=reg_loss_init.reshape(()), reg_loss=reg_loss.reshape(()),
                centerness_loss=centerness_loss.reshape(()))
