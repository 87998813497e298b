"""numpy restatement of the reference's IoU + label assignment -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows (paths relative to /root/reference/):
  pairwise_iou ...... detectron2/detectron2/structures/boxes.py:316-348 (area :173-182)
  Matcher ........... detectron2/detectron2/modeling/matcher.py:61-126
  TopKMatcher ....... slender_det/modeling/matchers/topk_matcher.py:38-86

Parity pin: checked in tests/test_oracle_assign.py against the reference's known-answer tests
(detectron2/tests/modeling/test_matcher.py:19-27, detectron2/tests/structures/test_boxes.py:151-173)
and against golden vectors produced by importing the reference's own Python files
(tests/golden/gen_golden.py).

All arithmetic is IEEE float32, one rounding per operation, in the reference's operation order,
so results are bit-comparable with the CUDA kernels (which use the non-fused _rn intrinsics).

Tie rule (TopKMatcher): the reference calls torch.topk, whose order among equal values is an
implementation detail of the backend.  The canonical rule used here and by the CUDA kernel is
"descending value, then ascending index" (== torch.sort(stable=True, descending=True)[:k]).
On inputs whose k-th and (k+1)-th largest values per GT row differ this equals the reference exactly.
"""
import numpy as np


def pairwise_iou(boxes1, boxes2):
    """boxes1 [N,4], boxes2 [M,4] (x1,y1,x2,y2) float32 -> IoU [N,M] float32."""
    b1 = np.asarray(boxes1, np.float32).reshape(-1, 4)
    b2 = np.asarray(boxes2, np.float32).reshape(-1, 4)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = np.minimum(b1[:, None, 2:], b2[None, :, 2:]) - np.maximum(b1[:, None, :2], b2[None, :, :2])
    wh = np.maximum(wh, np.float32(0))
    inter = wh[..., 0] * wh[..., 1]
    union = (area1[:, None] + area2[None, :]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(inter > 0, inter / union, np.float32(0))
    return iou.astype(np.float32)


def _threshold_labels(vals, thresholds, labels):
    th = [-np.inf] + [float(t) for t in thresholds] + [np.inf]
    assert len(labels) == len(th) - 1
    out = np.ones(vals.shape, np.int8)
    for l, lo, hi in zip(labels, th[:-1], th[1:]):
        out[(vals >= np.float32(lo)) & (vals < np.float32(hi))] = l
    return out


def matcher(q, thresholds, labels, allow_low_quality_matches=False):
    """q [M,N] float32 -> (matches int64 [N], labels int8 [N]).  d2 Matcher.__call__."""
    q = np.asarray(q, np.float32)
    assert q.ndim == 2
    M, N = q.shape
    if q.size == 0:
        return np.zeros(N, np.int64), np.full(N, labels[0], np.int8)
    assert (q >= 0).all()
    matches = q.argmax(axis=0).astype(np.int64)  # first (lowest GT index) on ties
    vals = q.max(axis=0)
    lab = _threshold_labels(vals, thresholds, labels)
    if allow_low_quality_matches:
        best = q.max(axis=1)
        lab[(q == best[:, None]).any(axis=0)] = 1
    return matches, lab


def topk_matcher(q, thresholds, labels, topk=9):
    """q [M,N] float32 -> (matches int64 [N], labels int8 [N]).  TopKMatcher.__call__."""
    q = np.asarray(q, np.float32)
    assert q.ndim == 2
    M, N = q.shape
    if q.size == 0:
        return np.zeros(N, np.int64), np.full(N, labels[0], np.int8)
    assert (q >= 0).all()
    if topk > N:
        raise RuntimeError("selected index k out of range")  # torch.topk's error
    matches = q.argmax(axis=0).astype(np.int64)
    vals = q.max(axis=0)
    lab = _threshold_labels(vals, thresholds, labels)
    idx = np.argsort(-q, axis=1, kind="stable")[:, :topk]
    lab[idx.reshape(-1)] = 1
    return matches, lab


def topk_is_tie_free(q, topk):
    """True when every GT row's k-th and (k+1)-th largest values differ (reference == canonical)."""
    q = np.asarray(q, np.float32)
    if q.size == 0 or q.shape[1] <= topk:
        return True
    s = -np.sort(-q, axis=1)
    return bool((s[:, topk - 1] != s[:, topk]).all())


def bbox_targets(candidates, gt, gt_labels, num_classes, pos_iou_thr=0.5, neg_iou_thr=0.4, gt_max_matching=True):
    """RepPointsV2.bbox_targets restated (reppointsv2.py:430-484), numpy.  `candidates` is clamped in place."""
    np.maximum(candidates, 0, out=candidates)                                  # :452-455
    overlaps = pairwise_iou(candidates, gt)                                    # [X, M]  :459
    labels = np.full((overlaps.shape[0],), num_classes, dtype=np.int64)        # :460
    max_ov, argmax_ov = overlaps.max(axis=1), overlaps.argmax(axis=1)          # :464
    gt_max = overlaps.max(axis=0)                                              # :467
    labels[max_ov < neg_iou_thr] = num_classes                                 # :469-470
    fg = max_ov >= pos_iou_thr                                                 # :472-473
    labels[fg] = gt_labels[argmax_ov[fg]]
    if gt_max_matching:                                                        # :475-477
        rows = np.nonzero(overlaps == gt_max[None, :])[0]
        labels[rows] = gt_labels[argmax_ov[rows]]
    boxes = np.zeros((overlaps.shape[0], 4), dtype=gt.dtype)                   # :479
    sel = (labels >= 0) & (labels != num_classes)                              # :481-482
    boxes[sel] = gt[argmax_ov[sel]]
    return boxes, labels


def point_targets(points, strides, gt, gt_labels, num_classes, scale=4):
    """RepPointsV2.point_targets restated (reppointsv2.py:370-428): the sequential loop over GTs, float32."""
    f = np.float32
    points = np.asarray(points, f)[:, :2]
    strides = np.asarray(strides, f)
    gt = np.asarray(gt, f)
    points_lvl = np.log2(strides).astype(np.int32)                              # :385
    lvl_min, lvl_max = points_lvl.min(), points_lvl.max()
    ctr = ((gt[:, :2] + gt[:, 2:]) / f(2)).astype(f)                            # :390
    wh = np.maximum(gt[:, 2:] - gt[:, :2], f(1e-6)).astype(f)                   # :391
    lvl = (((np.log2((wh[:, 0] / f(scale)).astype(f)) + np.log2((wh[:, 1] / f(scale)).astype(f))).astype(f) / f(2))
           .astype(f)).astype(np.int32)                                        # :395-396 (.int() truncates)
    lvl = np.clip(lvl, lvl_min, lvl_max)
    assigned = np.zeros((points.shape[0],), np.int64)
    dist_rec = np.full((points.shape[0],), np.inf, f)
    for idx in range(gt.shape[0]):                                              # :403-417
        sel = np.nonzero(points_lvl == lvl[idx])[0]
        if sel.size == 0:
            continue
        d = ((points[sel] - ctr[idx]) / wh[idx]).astype(f)
        dist = np.sqrt((d[:, 0] * d[:, 0]).astype(f) + (d[:, 1] * d[:, 1]).astype(f)).astype(f)
        j = int(np.argmin(dist))                                                # lowest index among equal minima
        if dist[j] < dist_rec[sel[j]]:
            assigned[sel[j]] = idx + 1
            dist_rec[sel[j]] = dist[j]
    boxes = np.zeros((points.shape[0], 4), gt.dtype)
    labels = np.full((points.shape[0],), num_classes, dtype=np.asarray(gt_labels).dtype)
    pos = assigned > 0
    labels[pos] = np.asarray(gt_labels)[assigned[pos] - 1]
    boxes[pos] = gt[assigned[pos] - 1]
    return boxes, labels


def fcos_location_targets(locations, soi, gt, gt_classes, num_points, strides, radius, num_classes, return_index=False):
    """compute_targets_for_locations for ONE image restated (fcos/utils.py:108-212), numpy float32."""
    f = np.float32
    INF = f(100000000)
    loc, soi, gt = np.asarray(locations, f), np.asarray(soi, f), np.asarray(gt, f)
    xs, ys = loc[:, 0], loc[:, 1]
    area = ((gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])).astype(f)
    l = xs[:, None] - gt[None, :, 0]; t = ys[:, None] - gt[None, :, 1]           # :176-180
    r = gt[None, :, 2] - xs[:, None]; b = gt[None, :, 3] - ys[:, None]
    reg = np.stack([l, t, r, b], axis=2).astype(f)
    if radius > 0:                                                               # get_sample_region :108-157
        cx = ((gt[:, 0] + gt[:, 2]) / f(2)).astype(f); cy = ((gt[:, 1] + gt[:, 3]) / f(2)).astype(f)
        if (cx[0] * len(xs)) == 0:                                               # :122-123 (sum of K copies)
            inside = np.zeros((len(xs), gt.shape[0]), bool)
        else:
            cg = np.zeros((len(xs), gt.shape[0], 4), f)
            beg = 0
            for level, n_p in enumerate(num_points):
                end = beg + n_p
                s = f(strides[level] * radius)
                xmin, ymin, xmax, ymax = cx - s, cy - s, cx + s, cy + s
                cg[beg:end, :, 0] = np.where(xmin > gt[:, 0], xmin, gt[:, 0])
                cg[beg:end, :, 1] = np.where(ymin > gt[:, 1], ymin, gt[:, 1])
                cg[beg:end, :, 2] = np.where(xmax > gt[:, 2], gt[:, 2], xmax)
                cg[beg:end, :, 3] = np.where(ymax > gt[:, 3], gt[:, 3], ymax)
                beg = end
            cb = np.stack([xs[:, None] - cg[..., 0], ys[:, None] - cg[..., 1], cg[..., 2] - xs[:, None],
                           cg[..., 3] - ys[:, None]], -1).astype(f)
            inside = cb.min(-1) > 0
    else:
        inside = reg.min(axis=2) > 0                                             # :186
    mx = reg.max(axis=2)
    cared = (mx >= soi[:, [0]]) & (mx <= soi[:, [1]])                            # :190-192
    a = np.repeat(area[None], len(xs), axis=0)
    a[~inside] = INF
    a[~cared] = INF
    idx = a.argmin(axis=1)                                                       # first minimum
    cls = np.asarray(gt_classes)[idx].copy()
    cls[a.min(axis=1) == INF] = num_classes                                      # :203
    if return_index:
        return cls, reg[np.arange(len(xs)), idx], idx
    return cls, reg[np.arange(len(xs)), idx]


def fcos_topk_locations(cls, reg, gt_index, num_classes, topk=5, slender=False):
    """The per-GT top-k-by-centerness selection of compute_topk_targets_for_locations (fcos/utils.py:264-279):
    `cls`, `reg` from fcos_location_targets, `gt_index` = the argmin GT of every location.
    slender=True: the FCOSRepPoints module's own copy of the loop (fcos_rpd_s1_topk.py:110-121), which scores with
    ITS compute_centerness_targets = pow(c, min(w/h, h/w)) (:25-55) instead of sqrt(c)."""
    f = np.float32
    fg = (cls >= 0) & (cls != num_classes)
    out = np.zeros((len(cls),), bool)
    for m in range(int(gt_index.max()) + 1 if len(gt_index) else 0):
        sel = np.nonzero((gt_index == m) & fg)[0]
        if sel.size > topk:
            r = reg[sel]
            c = ((np.minimum(r[:, 0], r[:, 2]) / np.maximum(r[:, 0], r[:, 2])).astype(f) *
                 (np.minimum(r[:, 1], r[:, 3]) / np.maximum(r[:, 1], r[:, 3])).astype(f)).astype(f)
            if slender:
                r1 = ((r[:, 0] + r[:, 2]).astype(f) / (r[:, 1] + r[:, 3]).astype(f)).astype(f)
                c = np.power(c, np.minimum(r1, (f(1) / r1).astype(f))).astype(f)
            else:
                c = np.sqrt(c).astype(f)
            order = np.lexsort((sel, -c))[:topk]                              # highest centerness, lowest index on ties
            out[sel[order]] = True
        elif sel.size > 0:
            out[sel] = True
    return out


def fcos_rpd_refine_targets(centers, init_boxes, gt, gt_classes, image_size, num_classes, thresholds=(0.4, 0.5),
                            labels=(0, -1, 1)):
    """FCOSRepPoints.get_ground_truth stage 2 restated (fcos_rpd_s1_topk.py:346-370), numpy float32."""
    f = np.float32
    centers, gt = np.asarray(centers, f), np.asarray(gt, f)
    q = pairwise_iou(gt, init_boxes)                                             # :352-354  [M, X]
    idx, matched = matcher(q, list(thresholds), list(labels), True)             # :355
    cls = np.asarray(gt_classes)[idx].copy()
    cls[matched == 0] = num_classes                                              # :357
    invalid = (centers[:, 0] >= image_size[1]) | (centers[:, 1] >= image_size[0])   # :349-350
    cls[invalid] = -1                                                            # :358
    box = gt[idx]
    xs, ys = centers[:, 0], centers[:, 1]
    reg = np.stack([xs - box[:, 0], ys - box[:, 1], box[:, 2] - xs, box[:, 3] - ys], axis=1).astype(f)   # :363-368
    return cls, reg
