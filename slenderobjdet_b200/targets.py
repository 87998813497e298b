"""Target assignment built on the fused IoU / matcher kernels.

``bbox_targets`` mirrors ``RepPointsV2.bbox_targets``
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:430-484): MaxIoU assignment of the
refined boxes.  The reference materialises the [X, M] IoU matrix, takes two maxima, and uses boolean-mask /
``nonzero`` indexing (host synchronisations); here the IoU + per-candidate argmax + per-GT maximum matching
is ONE pass of ``sdb_iou_assign`` and the gathers are ``torch.where`` - no IoU matrix, no host sync.
"""
import torch

from .matchers import Matcher, _tensor_of


@torch.no_grad()
def bbox_targets(candidate_bboxes, gt_bboxes, gt_labels, num_classes, pos_iou_thr=0.5, neg_iou_thr=0.4,
                 gt_max_matching=True):
    """-> (assigned_bboxes [X, 4], assigned_labels int64 [X]).

    * ``candidate_bboxes`` [X, 4] is clamped to >= 0 IN PLACE, as the reference does (:452-455);
    * a candidate is foreground when its best IoU is >= ``pos_iou_thr`` or (``gt_max_matching``) when its IoU
      with some GT equals that GT's maximum over all candidates (:471-477 - including GTs whose maximum is 0);
    * foreground rows get the box / label of their argmax GT, the rest ``num_classes`` and a zero box.  The
      ``neg_iou_thr`` band never changes the result (the labels start at ``num_classes``, :460, :468-469).
    """
    gt = _tensor_of(gt_bboxes)
    if candidate_bboxes.size(0) == 0 or gt.size(0) == 0:
        raise ValueError("No gt or anchors")
    candidate_bboxes.clamp_(min=0)
    matcher = Matcher([pos_iou_thr], [0, 1], allow_low_quality_matches=bool(gt_max_matching))
    matches, mlabels = matcher.from_boxes(gt, candidate_bboxes)
    fg = mlabels == 1
    labels = torch.where(fg, gt_labels.to(torch.long)[matches], torch.full_like(matches, num_classes))
    boxes = torch.where(fg[:, None], gt.to(candidate_bboxes.dtype)[matches], torch.zeros_like(candidate_bboxes))
    return boxes, labels


@torch.no_grad()
def point_targets(points, pts_strides, gt_bboxes, gt_labels, num_classes, point_base_scale=4):
    """``RepPointsV2.point_targets`` (reppointsv2.py:370-428): every GT claims the nearest point (normalised L2)
    of its pyramid level; contested points keep the closest GT (lowest GT index on ties).
    ``points`` [X, 2] or [X, >=2] (only the first two columns are used, as in the reference's callers),
    ``pts_strides`` [X], ``gt_bboxes`` [M, 4] xyxy, ``gt_labels`` [M] -> (assigned_bboxes [X, 4],
    assigned_labels [X]).  The reference's host loop over GTs (~12 launches + index ops per GT) becomes three
    launches."""
    from . import _lib
    gt = _tensor_of(gt_bboxes)
    if points.shape[0] == 0 or gt.shape[0] == 0:
        raise ValueError("No gt or bboxes")
    if not points.is_cuda:
        raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
    pts = points[:, :2].detach().float().contiguous()
    st = pts_strides.detach().float().contiguous()
    g = gt.detach().float().contiguous()
    gl = gt_labels.detach().to(torch.long).contiguous()
    X, M = pts.shape[0], g.shape[0]
    boxes = torch.empty((X, 4), dtype=torch.float32, device=pts.device)
    labels = torch.empty((X,), dtype=torch.long, device=pts.device)
    lib = _lib.lib()
    wsb = int(lib.sdb_point_targets_workspace_bytes(X))
    ws = torch.empty(wsb, dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(lib.sdb_point_targets(_lib.ptr(pts), _lib.ptr(st), _lib.ptr(g), _lib.ptr(gl), X, M,
                                         float(point_base_scale), int(num_classes), _lib.ptr(boxes), _lib.ptr(labels),
                                         _lib.ptr(ws), wsb, _lib.stream_ptr(pts.device)))
    return boxes.to(gt.dtype), labels.to(gt_labels.dtype)


def _boxes_and_classes(t):
    """Instances-like (``.gt_boxes`` / ``.gt_classes``) or a (boxes, classes) pair."""
    if hasattr(t, "gt_boxes"):
        return _tensor_of(t.gt_boxes), t.gt_classes
    b, c = t
    return _tensor_of(b), c


def _fcos_targets_batched(locations, targets, object_sizes_of_interest, strides, center_sampling_radius, num_classes,
                          topk, centerness_kind):
    """One ``sdb_fcos_location_targets_batched`` call for the whole batch: the GT boxes / classes of all images are
    padded to [N, M_pad] (one ``pad_sequence`` each), the per-image counts go along as a device int32 vector, and the
    kernels take the image as a grid dimension - no per-image host loop, two launches per batch."""
    import ctypes
    from torch.nn.utils.rnn import pad_sequence
    from . import _lib
    num_points = [len(l) for l in locations]
    loc = torch.cat(locations, dim=0).float().contiguous()
    if not loc.is_cuda:
        raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
    soi = object_sizes_of_interest.float().contiguous()
    X, L, N = loc.shape[0], len(num_points), len(targets)
    pairs = [_boxes_and_classes(t) for t in targets]
    counts = [int(b.shape[0]) for b, _ in pairs]
    if min(counts) == 0:
        raise ValueError("every image needs at least one GT box (the reference indexes an empty tensor otherwise)")
    gt = pad_sequence([b.float() for b, _ in pairs], batch_first=True).contiguous()              # [N, M_pad, 4]
    gc = pad_sequence([c.to(torch.long) for _, c in pairs], batch_first=True).contiguous()        # [N, M_pad]
    cnt = torch.tensor(counts, dtype=torch.int32).to(loc.device, non_blocking=True)
    npl = (ctypes.c_int32 * L)(*num_points)
    lst = (ctypes.c_float * L)(*[float(s) for s in strides])
    lib = _lib.lib()
    oc = torch.empty((N, X), dtype=torch.long, device=loc.device)
    orr = torch.empty((N, X, 4), dtype=torch.float32, device=loc.device)
    ot = torch.empty((N, X), dtype=torch.uint8, device=loc.device) if topk else None
    wsb = N * int(lib.sdb_fcos_topk_workspace_bytes(X)) if topk else 0
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=loc.device)
    with torch.cuda.device(loc.device):
        _lib.check(lib.sdb_fcos_location_targets_batched(
            _lib.ptr(loc), _lib.ptr(soi), _lib.ptr(gt), _lib.ptr(gc), _lib.ptr(cnt), N, X, gt.shape[1], npl, lst, L,
            float(center_sampling_radius), int(num_classes), int(topk), int(centerness_kind), _lib.ptr(oc), _lib.ptr(orr),
            _lib.ptr(ot), _lib.ptr(ws), wsb, _lib.stream_ptr(loc.device)))
    return oc.to(pairs[0][1].dtype), orr.to(pairs[0][0].dtype), (ot.bool() if topk else None), num_points


@torch.no_grad()
def compute_targets_for_locations(locations, targets, object_sizes_of_interest, strides, center_sampling_radius,
                                  num_classes):
    """``slender_det.modeling.meta_arch.fcos.utils.compute_targets_for_locations`` (fcos/utils.py:160-212).

    ``locations``: list of per-level [X_l, 2] tensors; ``targets``: per image an Instances-like object or a
    (boxes [M,4], classes [M]) pair; ``object_sizes_of_interest`` [X, 2]; ``strides``: per-level ints.
    -> (gt_classes [N, X], reg_targets [N, X, 4]).  ONE fused kernel for the whole batch (image = grid dimension,
    GT boxes in shared memory, one thread per location) instead of [X, M, 4] temporaries and ~25 launches per
    image; results are identical."""
    cls, reg, _, _ = _fcos_targets_batched(locations, targets, object_sizes_of_interest, strides,
                                           center_sampling_radius, num_classes, 0, 0)
    return cls, reg


@torch.no_grad()
def compute_topk_targets_for_locations(locations, targets, object_sizes_of_interest, strides, center_sampling_radius,
                                       num_classes, norm_reg_targets=False, topk=5, slender_centerness=False):
    """``compute_topk_targets_for_locations`` (fcos/utils.py:215-292): ``compute_targets_for_locations`` plus, per GT,
    its ``topk`` foreground locations of highest centerness.  ``slender_centerness=True`` scores with the
    FCOSRepPoints module's own pow-form centerness, as ITS copy of this function does (fcos_rpd_s1_topk.py:57-134).
    -> (gt_classes [N, X], reg_targets [N, X, 4], topk_locations bool [N, X]).  Two launches per BATCH; the
    reference's host loops over images and GTs (a ``.sum().item()`` sync per GT) are gone."""
    cls, reg, tk, num_points = _fcos_targets_batched(locations, targets, object_sizes_of_interest, strides,
                                                     center_sampling_radius, num_classes, topk,
                                                     1 if slender_centerness else 0)
    if norm_reg_targets:   # :221-223, :284-285 (built on the CPU in the reference, on the locations' device here)
        norm_weights = torch.cat([torch.empty(n).fill_(s) for n, s in zip(num_points, strides)]).to(reg.device)
        reg = reg / norm_weights[None, :, None]
    return cls, reg, tk


@torch.no_grad()
def fcos_rpd_refine_targets(centers, init_boxes, gt_boxes, gt_classes, image_size, num_classes,
                            iou_thresholds=(0.4, 0.5), iou_labels=(0, -1, 1), allow_low_quality_matches=True):
    """Stage 2 of ``FCOSRepPoints.get_ground_truth`` for one image (fcos_rpd_s1_topk.py:346-370): match the
    stage-1 boxes to the GTs by IoU (``Matcher(RETINANET.IOU_THRESHOLDS, IOU_LABELS, allow_low_quality=True)``,
    :174-178), class labels (``num_classes`` where the match label is 0, GT class where it is 1 or -1, then -1
    for centres outside the image) and ltrb refine targets to the matched GT.
    ``centers`` [X,2], ``init_boxes`` [X,4], ``gt_boxes`` [M,4], ``gt_classes`` [M], ``image_size`` (h, w).
    One fused IoU + matcher pass instead of the [M, X] matrix; no host sync."""
    gt = _tensor_of(gt_boxes)
    matcher = Matcher(list(iou_thresholds), list(iou_labels), allow_low_quality_matches=allow_low_quality_matches)
    idx, matched = matcher.from_boxes(gt, _tensor_of(init_boxes))
    cls = gt_classes[idx]
    cls = torch.where(matched == 0, torch.full_like(cls, num_classes), cls)
    invalid = (centers[:, 0] >= image_size[1]).logical_or(centers[:, 1] >= image_size[0])
    cls = torch.where(invalid, torch.full_like(cls, -1), cls)
    box = gt[idx]
    xs, ys = centers[:, 0], centers[:, 1]
    reg = torch.stack([xs - box[:, 0], ys - box[:, 1], box[:, 2] - xs, box[:, 3] - ys], dim=1)
    return cls, reg


@torch.no_grad()
def fcos_rpd_get_ground_truth(points, init_boxes, gt_instances, fpn_strides, center_sampling_radius, num_classes,
                              iou_thresholds=(0.4, 0.5), iou_labels=(0, -1, 1), topk=5):
    """``FCOSRepPoints.get_ground_truth`` (fcos_rpd_s1_topk.py:320-376): stage-1 FCOS location targets with the
    per-GT top-k-by-centerness mask, and stage-2 IoU-matched refine targets.  ``gt_instances``: per image an
    Instances-like object (``gt_boxes``, ``gt_classes``, ``image_size``) or a (boxes, classes, image_size) triple.
    -> (init_gt_classes, init_reg_targets, refine_gt_classes, refine_reg_targets, topk_locations)."""
    INF = 100000000
    sizes = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, INF]]
    soi = torch.cat([p.new_tensor(sizes[l])[None].expand(len(p), -1) for l, p in enumerate(points)], dim=0)
    trip = []
    for t in gt_instances:
        if hasattr(t, "gt_boxes"):
            trip.append((_tensor_of(t.gt_boxes), t.gt_classes, t.image_size))
        else:
            trip.append((_tensor_of(t[0]), t[1], t[2]))
    init_cls, init_reg, topk_loc = compute_topk_targets_for_locations(
        points, [(b, c) for b, c, _ in trip], soi, fpn_strides, center_sampling_radius, num_classes, topk=topk,
        slender_centerness=True)   # the module's own pow-form centerness ranks the top-5 (:25-55, :117)
    centers = torch.cat(points, 0)
    cls_labels, reg_labels = [], []
    for i, (b, c, sz) in enumerate(trip):
        cl, rg = fcos_rpd_refine_targets(centers, init_boxes[i], b, c, sz, num_classes, iou_thresholds, iou_labels)
        cls_labels.append(cl)
        reg_labels.append(rg)
    return init_cls, init_reg, torch.stack(cls_labels), torch.stack(reg_labels), topk_loc
