"""CPU restatement of the dense point head's inference post-processing -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows RepPointsV2.inference_single_image
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:533-603) and pts_to_bbox (:328-366):
per level sigmoid -> sort descending -> first `topk` -> score > threshold -> decode the surviving points' boxes
(point set -> box, x stride, + centre, clamped to the image) -> concatenate levels -> class-aware NMS
(detectron2/layers/nms.py:10-29 -> torchvision batched_nms: greedy, IoU > threshold suppresses, same class only)
-> first `max_det` in decreasing score order.

Pinned on tests/golden/postproc_cases.npz, produced by executing the reference's inference_single_image by file path
(tests/golden/gen_postproc_golden.py).  float32 arithmetic in the reference's operation order; ties between equal
scores are broken by the lower flat index (the reference's torch.sort is not stable: implementation-defined there).
"""
import numpy as np

F = np.float32


def pts_to_bbox(pts, transform="minmax", moment_transfer=None):
    """pts [P, 2n] (x0, y0, x1, y1, ...) -> [P, 4]  (reppointsv2.py:328-366)."""
    x, y = pts[:, 0::2], pts[:, 1::2]
    if transform == "minmax":
        return np.stack([x.min(1), y.min(1), x.max(1), y.max(1)], 1).astype(F)
    if transform == "partial_minmax":
        return np.stack([x[:, :4].min(1), y[:, :4].min(1), x[:, :4].max(1), y[:, :4].max(1)], 1).astype(F)
    if transform == "moment":
        mx, my = x.mean(1, dtype=F), y.mean(1, dtype=F)
        sx, sy = x.std(1, ddof=1, dtype=F), y.std(1, ddof=1, dtype=F)      # torch.std: unbiased
        hw, hh = sx * F(np.exp(F(moment_transfer[0]))), sy * F(np.exp(F(moment_transfer[1])))
        return np.stack([mx - hw, my - hh, mx + hw, my + hh], 1).astype(F)
    raise ValueError(transform)


def nms(boxes, scores, classes, thr):
    """Greedy class-aware NMS; returns kept indices in decreasing score order."""
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    b = boxes.astype(F)
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(F)
    removed = np.zeros(len(scores), bool)
    keep = []
    for i in order:
        if removed[i]:
            continue
        keep.append(i)
        w = np.maximum(np.minimum(b[i, 2], b[:, 2]) - np.maximum(b[i, 0], b[:, 0]), F(0)).astype(F)
        h = np.maximum(np.minimum(b[i, 3], b[:, 3]) - np.maximum(b[i, 1], b[:, 1]), F(0)).astype(F)
        inter = (w * h).astype(F)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = (inter / ((area[i] + area).astype(F) - inter).astype(F)).astype(F)
        removed |= (iou > F(thr)) & (classes == classes[i])
        removed[i] = True
    return np.asarray(keep, np.int64)


def inference_single_image(cls_logits, pts_refine, strides, points, image_size, num_classes, score_thresh=0.05,
                           topk=1000, nms_thresh=0.5, max_det=100, transform="minmax", moment_transfer=None):
    """cls_logits[l] [HW, K], pts_refine[l] [HW, 2n], strides[l] scalar, points[l] [HW, 2], image_size (h, w).
    -> boxes [D, 4] float32, scores [D] float32, classes [D] int64 with D <= max_det."""
    boxes_all, scores_all, cls_all = [], [], []
    for logits, pts, stride, ctr in zip(cls_logits, pts_refine, strides, points):
        box = pts_to_bbox(np.asarray(pts, F), transform, moment_transfer)
        box = (box * F(stride) + np.concatenate([ctr, ctr], 1).astype(F)).astype(F)                # :558-560
        box[:, 0::2] = np.clip(box[:, 0::2], F(0), F(image_size[1]))                                # :561-564
        box[:, 1::2] = np.clip(box[:, 1::2], F(0), F(image_size[0]))
        prob = (F(1) / (F(1) + np.exp(-np.asarray(logits, F).reshape(-1)))).astype(F)               # :567
        n = min(topk, prob.size)                                                                    # :570
        order = np.lexsort((np.arange(prob.size), -prob.astype(np.float64)))[:n]                    # :572-574
        order = order[prob[order] > F(score_thresh)]                                                # :577-579
        boxes_all.append(box[order // num_classes])                                                 # :581-584
        scores_all.append(prob[order])
        cls_all.append((order % num_classes).astype(np.int64))
    b, s, c = np.concatenate(boxes_all), np.concatenate(scores_all), np.concatenate(cls_all)
    keep = nms(b, s, c, nms_thresh)[:max_det]                                                       # :595-596
    return b[keep], s[keep], c[keep]
