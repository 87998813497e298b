"""CPU: pin the numpy IoU/matcher oracle against upstream known answers and reference goldens."""
import numpy as np
import pytest

from oracle import assign as oa


def test_known_answer_pairwise_iou(assign_cases):
    """detectron2/tests/structures/test_boxes.py:151-173"""
    c = assign_cases["ka_iou"]
    assert np.allclose(oa.pairwise_iou(c["boxes1"], c["boxes2"]), c["expected"])


def test_known_answer_matcher(assign_cases):
    """detectron2/tests/modeling/test_matcher.py:19-27"""
    c = assign_cases["ka_matcher"]
    m, l = oa.matcher(c["q"], list(c["thresholds"]), list(c["label_values"]), allow_low_quality_matches=True)
    assert np.array_equal(m, c["matches"]) and np.array_equal(l, c["labels"])
    assert m.dtype == np.int64 and l.dtype == np.int8


@pytest.mark.parametrize("name", ["small", "mid", "slender"])
def test_iou_bit_exact_vs_reference(assign_cases, name):
    c = assign_cases[name]
    iou = oa.pairwise_iou(c["gt"], c["anchors"])
    assert iou.dtype == np.float32
    assert np.array_equal(iou.view(np.uint32), c["iou"].view(np.uint32))


@pytest.mark.parametrize("name", ["small", "mid", "slender"])
@pytest.mark.parametrize("k", [1, 9, 10])
def test_topk_matcher_vs_reference(assign_cases, name, k):
    c = assign_cases[name]
    m, l = oa.topk_matcher(c["iou"], [0.3, 0.7], [0, -1, 1], topk=k)
    assert np.array_equal(m, c[f"topk{k}_matches"])
    assert bool(c[f"topk{k}_tiefree"]) == oa.topk_is_tie_free(c["iou"], k)
    if bool(c[f"topk{k}_tiefree"]):
        assert np.array_equal(l, c[f"topk{k}_labels"])
    else:  # canonical rule may differ from torch.topk only in WHICH tied anchors get label 1
        assert (l == 1).sum() == (c[f"topk{k}_labels"] == 1).sum()


@pytest.mark.parametrize("name", ["small", "mid", "slender"])
@pytest.mark.parametrize("alq", [0, 1])
def test_matcher_vs_reference(assign_cases, name, alq):
    c = assign_cases[name]
    m, l = oa.matcher(c["iou"], [0.4, 0.5], [0, -1, 1], allow_low_quality_matches=bool(alq))
    assert np.array_equal(m, c[f"matcher{alq}_matches"]) and np.array_equal(l, c[f"matcher{alq}_labels"])


def test_empty_gt(assign_cases):
    c = assign_cases["empty"]
    for fn in (oa.topk_matcher, oa.matcher):
        m, l = fn(np.zeros((0, 11), np.float32), [0.3, 0.7], [0, -1, 1])
        assert np.array_equal(m, c["matches"]) and np.array_equal(l, c["labels"])
        assert m.dtype == np.int64 and l.dtype == np.int8


def test_tie_rule_is_lowest_index():
    q = np.zeros((2, 12), np.float32)
    q[0, 5] = 0.9
    q[1, 3] = 0.2
    m, l = oa.topk_matcher(q, [0.3, 0.7], [0, -1, 1], topk=3)
    # GT0: anchor 5 then lowest-index zeros {0,1}; GT1: anchor 3 then {0,1}
    assert sorted(np.nonzero(l == 1)[0].tolist()) == [0, 1, 3, 5]
    assert m[5] == 0 and m[3] == 1 and m[0] == 0


def _bbox_targets_reference_lines(candidate_bboxes, gt, gt_labels, num_classes, pos_iou_thr=0.5, neg_iou_thr=0.4,
                                  gt_max_matching=True):
    """reppointsv2.py:452-484 in torch (CPU), with the IoU of boxes.py:333-347."""
    import torch
    candidate_bboxes[:, 0].clamp_(min=0); candidate_bboxes[:, 1].clamp_(min=0)
    candidate_bboxes[:, 2].clamp_(min=0); candidate_bboxes[:, 3].clamp_(min=0)
    b1, b2 = candidate_bboxes, gt
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    overlaps = torch.where(inter > 0, inter / (area1[:, None] + area2 - inter), torch.zeros(1, dtype=inter.dtype))
    assigned_labels = overlaps.new_full((overlaps.size(0),), num_classes, dtype=torch.long)
    max_overlaps, argmax_overlaps = overlaps.max(dim=1)
    gt_max_overlaps, _ = overlaps.max(dim=0)
    assigned_labels[max_overlaps < neg_iou_thr] = num_classes
    fg_inds = max_overlaps >= pos_iou_thr
    assigned_labels[fg_inds] = gt_labels[argmax_overlaps[fg_inds]]
    if gt_max_matching:
        fg_inds = torch.nonzero(overlaps == gt_max_overlaps)[:, 0]
        assigned_labels[fg_inds] = gt_labels[argmax_overlaps[fg_inds]]
    assigned_bboxes = overlaps.new_zeros((b1.size(0), 4))
    fg_inds = (assigned_labels >= 0) & (assigned_labels != num_classes)
    assigned_bboxes[fg_inds] = gt[argmax_overlaps[fg_inds]]
    return assigned_bboxes, assigned_labels


def _bbox_case(seed, X=3000, M=23):
    import torch
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(X, 2, generator=g) * torch.tensor([640.0, 480.0])
    wh = torch.exp(torch.rand(X, 2, generator=g) * 3.5 + 1.5)
    cand = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)            # some coordinates negative: exercises the clamp
    c2 = torch.rand(M, 2, generator=g) * torch.tensor([640.0, 480.0])
    wh2 = torch.exp(torch.rand(M, 2, generator=g) * 3.5 + 2.0)
    gt = torch.cat([c2 - wh2 / 2, c2 + wh2 / 2], 1).clamp(min=0)
    gt[0] = torch.tensor([5000.0, 5000.0, 5100.0, 5100.0])      # a GT nothing overlaps: its maximum is 0
    labels = torch.randint(0, 80, (M,), generator=g)
    return cand, gt, labels


def test_bbox_targets_oracle_matches_reference_lines():
    for seed, gmm in ((0, True), (1, False), (2, True)):
        cand, gt, labels = _bbox_case(seed)
        c1, c2 = cand.clone(), cand.clone().numpy()
        rb, rl = _bbox_targets_reference_lines(c1, gt, labels, 80, gt_max_matching=gmm)
        ob, ol_ = oa.bbox_targets(c2, gt.numpy(), labels.numpy(), 80, gt_max_matching=gmm)
        assert np.array_equal(c1.numpy(), c2)                     # same in-place clamp
        assert np.array_equal(ol_, rl.numpy()) and np.array_equal(ob, rb.numpy())


def _point_targets_reference_lines(points, pts_strides, gt_bboxes, gt_labels, num_classes, point_base_scale=4):
    """reppointsv2.py:383-428 in torch (CPU)."""
    import torch
    points_lvl = torch.log2(pts_strides).int()
    lvl_min, lvl_max = points_lvl.min(), points_lvl.max()
    num_gts, num_points = gt_bboxes.shape[0], points.shape[0]
    gt_bboxes_ctr_xy = (gt_bboxes[:, :2] + gt_bboxes[:, 2:]) / 2
    gt_bboxes_wh = (gt_bboxes[:, 2:] - gt_bboxes[:, :2]).clamp(min=1e-6)
    scale = point_base_scale
    gt_bboxes_lvl = ((torch.log2(gt_bboxes_wh[:, 0] / scale) + torch.log2(gt_bboxes_wh[:, 1] / scale)) / 2).int()
    gt_bboxes_lvl = torch.clamp(gt_bboxes_lvl, min=lvl_min, max=lvl_max)
    assigned_gt_inds = points.new_zeros((num_points,), dtype=torch.long)
    assigned_gt_dist = points.new_full((num_points,), float('inf'))
    points_range = torch.arange(points.shape[0])
    for idx in range(num_gts):
        gt_lvl = gt_bboxes_lvl[idx]
        lvl_idx = gt_lvl == points_lvl
        points_index = points_range[lvl_idx]
        lvl_points = points[lvl_idx, :]
        gt_point = gt_bboxes_ctr_xy[[idx], :]
        gt_wh = gt_bboxes_wh[[idx], :]
        points_gt_dist = ((lvl_points - gt_point) / gt_wh).norm(dim=1)
        min_dist, min_dist_index = torch.topk(points_gt_dist, 1, largest=False)
        min_dist_points_index = points_index[min_dist_index]
        less_than_recorded_index = min_dist < assigned_gt_dist[min_dist_points_index]
        min_dist_points_index = min_dist_points_index[less_than_recorded_index]
        assigned_gt_inds[min_dist_points_index] = idx + 1
        assigned_gt_dist[min_dist_points_index] = min_dist[less_than_recorded_index]
    assigned_bboxes = gt_bboxes.new_zeros((num_points, 4))
    assigned_labels = gt_labels.new_full((num_points,), num_classes)
    pos_inds = torch.nonzero(assigned_gt_inds > 0).squeeze().long()
    if pos_inds.numel() > 0:
        assigned_labels[pos_inds] = gt_labels[assigned_gt_inds[pos_inds] - 1]
        assigned_bboxes[pos_inds] = gt_bboxes[assigned_gt_inds[pos_inds] - 1]
    return assigned_bboxes, assigned_labels


def _points_case(seed, M=60, levels=((25, 42, 8), (13, 21, 16), (7, 11, 32), (4, 6, 64), (2, 3, 128))):
    import torch
    g = torch.Generator().manual_seed(seed)
    pts, strides = [], []
    for (h, w, s) in levels:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) * s, ys.reshape(-1) * s], 1))      # reference: shifts = arange * stride
        strides.append(torch.full((h * w,), float(s)))
    pts, strides = torch.cat(pts), torch.cat(strides)
    c = torch.rand(M, 2, generator=g) * torch.tensor([320.0, 190.0])
    wh = torch.exp(torch.rand(M, 2, generator=g) * 5.0 + 0.5)                     # 1.6 .. 245 px: every level + clamping
    gt = torch.cat([c - wh / 2, c + wh / 2], 1)
    gt[1] = gt[0]                                                                 # two GTs claim the same point at equal distance
    labels = torch.randint(0, 80, (M,), generator=g)
    return pts, strides, gt, labels


def test_point_targets_oracle_matches_reference_lines():
    for seed in (0, 1, 2):
        pts, strides, gt, labels = _points_case(seed)
        rb, rl = _point_targets_reference_lines(pts, strides, gt, labels, 80)
        ob, ol_ = oa.point_targets(pts.numpy(), strides.numpy(), gt.numpy(), labels.numpy(), 80)
        assert np.array_equal(ol_, rl.numpy()) and np.array_equal(ob, rb.numpy())
        assert (rl != 80).sum() > 10


def _fcos_reference_lines(locations, boxes, classes, soi, strides, radius, num_classes):
    """fcos/utils.py:108-212 for one image, in torch (CPU), with the reference's own helper restated verbatim."""
    import torch
    INF = 100000000
    num_points = [len(_) for _ in locations]
    locations = torch.cat(locations, dim=0)
    xs, ys = locations[:, 0], locations[:, 1]
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    l = xs[:, None] - boxes[:, 0][None]; t = ys[:, None] - boxes[:, 1][None]
    r = boxes[:, 2][None] - xs[:, None]; b = boxes[:, 3][None] - ys[:, None]
    reg = torch.stack([l, t, r, b], dim=2)
    if radius > 0:
        gt = boxes; K = len(xs); num_gts = gt.shape[0]
        gt = gt[None].expand(K, num_gts, 4)
        center_x = (gt[..., 0] + gt[..., 2]) / 2
        center_y = (gt[..., 1] + gt[..., 3]) / 2
        center_gt = gt.new_zeros(gt.shape)
        if center_x[..., 0].sum() == 0:
            is_in = xs.new_zeros(xs.shape, dtype=torch.uint8)[:, None].expand(K, num_gts)
        else:
            beg = 0
            for level, n_p in enumerate(num_points):
                end = beg + n_p
                stride = strides[level] * radius
                xmin = center_x[beg:end] - stride; ymin = center_y[beg:end] - stride
                xmax = center_x[beg:end] + stride; ymax = center_y[beg:end] + stride
                center_gt[beg:end, :, 0] = torch.where(xmin > gt[beg:end, :, 0], xmin, gt[beg:end, :, 0])
                center_gt[beg:end, :, 1] = torch.where(ymin > gt[beg:end, :, 1], ymin, gt[beg:end, :, 1])
                center_gt[beg:end, :, 2] = torch.where(xmax > gt[beg:end, :, 2], gt[beg:end, :, 2], xmax)
                center_gt[beg:end, :, 3] = torch.where(ymax > gt[beg:end, :, 3], gt[beg:end, :, 3], ymax)
                beg = end
            cb = torch.stack((xs[:, None] - center_gt[..., 0], ys[:, None] - center_gt[..., 1],
                              center_gt[..., 2] - xs[:, None], center_gt[..., 3] - ys[:, None]), -1)
            is_in = cb.min(-1)[0] > 0
    else:
        is_in = reg.min(dim=2)[0] > 0
    mx = reg.max(dim=2)[0]
    cared = (mx >= soi[:, [0]]) & (mx <= soi[:, [1]])
    a = area[None].repeat(len(locations), 1)
    a[is_in == 0] = INF
    a[cared == 0] = INF
    mn, ind = a.min(dim=1)
    cls = classes[ind]
    cls[mn == INF] = num_classes
    return cls, reg[range(len(locations)), ind]


def _fcos_case(seed, M=40, levels=((25, 42, 8), (13, 21, 16), (7, 11, 32), (4, 6, 64), (2, 3, 128))):
    import torch
    g = torch.Generator().manual_seed(seed)
    locs, soi = [], []
    ranges = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, 100000000]]
    for (h, w, s), rg in zip(levels, ranges):
        ys, xs = torch.meshgrid(torch.arange(0, h * s, s, dtype=torch.float32), torch.arange(0, w * s, s, dtype=torch.float32),
                                indexing="ij")
        locs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), dim=1) + s // 2)     # compute_locations_per_level
        soi.append(torch.tensor(rg, dtype=torch.float32)[None].expand(h * w, -1))
    W, H = levels[0][1] * levels[0][2], levels[0][0] * levels[0][2]
    c = torch.rand(M, 2, generator=g) * torch.tensor([float(W), float(H)])
    wh = torch.exp(torch.rand(M, 2, generator=g) * 4.5 + 1.5)
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(min=0)
    boxes[3] = boxes[2]                                                                # equal areas: first index wins
    classes = torch.randint(0, 80, (M,), generator=g)
    return locs, torch.cat(soi), boxes, classes, [l[2] for l in levels]


def test_fcos_location_targets_oracle_matches_reference_lines():
    for seed, radius in ((0, 0.0), (1, 1.5), (2, 1.0)):
        locs, soi, boxes, classes, strides = _fcos_case(seed)
        rc, rr = _fcos_reference_lines(locs, boxes, classes, soi, strides, radius, 80)
        import torch
        oc, orr = oa.fcos_location_targets(torch.cat(locs).numpy(), soi.numpy(), boxes.numpy(), classes.numpy(),
                                           [len(l) for l in locs], strides, radius, 80)
        assert np.array_equal(oc, rc.numpy()) and np.array_equal(orr, rr.numpy())
        assert (rc != 80).sum() > 20


def test_fcos_topk_oracle_matches_reference_loop():
    """fcos/utils.py:264-279 in torch on CPU against the oracle's selection (tie-free centerness)."""
    import torch
    locs, soi, boxes, classes, strides = _fcos_case(3, 12)
    cls, reg, idx = oa.fcos_location_targets(torch.cat(locs).numpy(), soi.numpy(), boxes.numpy(), classes.numpy(),
                                             [len(l) for l in locs], strides, 0.0, 80, return_index=True)
    got = oa.fcos_topk_locations(cls, reg, idx, 80, topk=5)
    gt_classes_per_im, reg_t, inds_t = torch.from_numpy(cls), torch.from_numpy(reg), torch.from_numpy(idx)
    fore = (gt_classes_per_im >= 0) & (gt_classes_per_im != 80)
    ref = torch.zeros(len(inds_t)).bool()
    for gi in range(len(boxes)):
        sel = (inds_t == gi) & fore
        n = sel.sum().item()
        if n > 5:
            r = reg_t[sel]
            lr, tb = r[:, [0, 2]], r[:, [1, 3]]
            score = torch.sqrt((lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0]))
            _, ii = torch.topk(score, 5, sorted=False)
            ref[sel.nonzero()[ii]] = True
        elif n > 0:
            ref[sel.nonzero()] = True
    assert np.array_equal(got, ref.numpy()) and 0 < got.sum() < fore.sum().item()


def test_fcos_rpd_refine_targets_oracle_matches_reference_lines():
    """fcos_rpd_s1_topk.py:346-370 with the reference's own Matcher / pairwise_iou (imported by path in
    gen_golden.py; restated here with torch on CPU)."""
    import torch
    locs, soi, boxes, classes, strides = _fcos_case(14, 30)
    centers = torch.cat(locs)
    g = torch.Generator().manual_seed(2)
    wh = torch.exp(torch.rand(centers.shape[0], 2, generator=g) * 3.0 + 2.0)
    init = torch.cat([centers - wh / 2, centers + wh / 2], 1)
    image_size = (160, 300)
    # pairwise_iou (boxes.py:333-347) + Matcher with allow_low_quality_matches (matcher.py:61-126)
    b1, b2 = boxes, init
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1]); a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    whi = (torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])).clamp(min=0)
    inter = whi.prod(dim=2)
    q = torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1))
    matched_vals, matches = q.max(dim=0)
    match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
    for l, low, high in zip([0, -1, 1], [-float("inf"), 0.4, 0.5], [0.4, 0.5, float("inf")]):
        match_labels[(matched_vals >= low) & (matched_vals < high)] = l
    highest, _ = q.max(dim=1)
    _, pred_inds = torch.nonzero(q == highest[:, None], as_tuple=True)
    match_labels[pred_inds] = 1
    cls_label = classes[matches]
    cls_label[match_labels == 0] = 80
    invalid = (centers[:, 0] >= image_size[1]).logical_or(centers[:, 1] >= image_size[0])
    cls_label[invalid] = -1
    rb = boxes[matches]
    xs, ys = centers[:, 0], centers[:, 1]
    reg = torch.stack([xs - rb[:, 0], ys - rb[:, 1], rb[:, 2] - xs, rb[:, 3] - ys], dim=1)
    oc, orr = oa.fcos_rpd_refine_targets(centers.numpy(), init.numpy(), boxes.numpy(), classes.numpy(), image_size, 80)
    assert np.array_equal(oc, cls_label.numpy()) and np.array_equal(orr, reg.numpy())
