"""Inference post-processing of a dense point head on the GPU, whole batch in one native call.

Replaces ``RepPointsV2.inference`` / ``inference_single_image``
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:486-603): per level sigmoid, top-k candidates
above the score threshold, box decoding from the refined point sets, then class-aware NMS over all levels and the
``max_detections`` best survivors.  The reference permutes the head outputs to ``[HW, K]``, sorts every level's full
score vector, indexes with boolean masks and runs ``batched_nms`` per image from a Python loop; here the head's NCHW
outputs are read in place and the whole batch is seven launches with no host synchronisation in between.
"""
import ctypes

import torch

from . import _lib

_TRANSFORMS = {"minmax": 0, "partial_minmax": 1, "moment": 2}


@torch.no_grad()
def points_inference(center_pts, cls_outs, pts_outs_refine, fpn_strides, image_sizes, num_classes, score_threshold=0.05,
                     topk_candidates=1000, nms_threshold=0.5, max_detections=100, transform_method="minmax",
                     moment_transfer=None):
    """center_pts: per level ``[H*W, 2]`` (x, y) point centres (the same for every image);
    cls_outs / pts_outs_refine: per level ``[N, K, H, W]`` / ``[N, 2*num_points, H, W]`` head outputs;
    fpn_strides: per-level stride; image_sizes: per image ``(height, width)``.
    -> ``(boxes [N, max_detections, 4], scores [N, max_detections], classes int64 [N, max_detections], counts int32 [N])``
    on the device, rows beyond ``counts[n]`` undefined.  ``split_detections`` turns them into per-image tensors (that step
    reads the counts back, as building the reference's ``Instances`` does)."""
    lib = _lib.lib()
    L, N = len(cls_outs), cls_outs[0].shape[0]
    dev = cls_outs[0].device
    if not cls_outs[0].is_cuda:
        raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
    if len(pts_outs_refine) != L or len(center_pts) != L or len(fpn_strides) != L or len(image_sizes) != N:
        raise ValueError("one entry per level / image expected")
    keep = []   # keep the contiguous float32 views alive until the call is enqueued
    rows = []
    num_points = pts_outs_refine[0].shape[1] // 2
    for l in range(L):
        c = cls_outs[l].detach().float().contiguous()
        p = pts_outs_refine[l].detach().float().contiguous()
        ctr = center_pts[l].detach().float().contiguous()
        if c.shape[1] != num_classes or p.shape[1] != 2 * num_points or tuple(c.shape[2:]) != tuple(p.shape[2:]) \
                or ctr.shape[0] != c.shape[2] * c.shape[3]:
            raise ValueError("level %d: inconsistent shapes" % l)
        keep += [c, p, ctr]
        rows.append(_lib.PPLevel(_lib.addr(c), _lib.addr(p), _lib.addr(ctr), c.shape[2], c.shape[3], float(fpn_strides[l])))
    levels = (_lib.PPLevel * L)(*rows)
    sizes = (ctypes.c_int32 * (2 * N))(*[int(v) for hw in image_sizes for v in hw])
    mt = None
    if transform_method == "moment":
        m = moment_transfer.detach().float().cpu() if torch.is_tensor(moment_transfer) else moment_transfer
        mt = (ctypes.c_float * 2)(float(m[0]), float(m[1]))
    boxes = torch.empty((N, max_detections, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((N, max_detections), dtype=torch.float32, device=dev)
    classes = torch.empty((N, max_detections), dtype=torch.int64, device=dev)
    counts = torch.empty((N,), dtype=torch.int32, device=dev)
    wsb = int(lib.sdb_points_postprocess_workspace_bytes(L, N, int(topk_candidates), float(score_threshold)))
    if wsb == 0:
        raise RuntimeError("slender_b200: unsupported post-processing configuration (levels <= 8, images <= 64, topk <= 2048)")
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sdb_points_postprocess(levels, L, N, int(num_classes), int(num_points), _TRANSFORMS[transform_method],
                                              mt, sizes, float(score_threshold), int(topk_candidates), float(nms_threshold),
                                              int(max_detections), _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(classes),
                                              _lib.ptr(counts), None, _lib.ptr(ws), wsb, _lib.stream_ptr(dev)))
    return boxes, scores, classes, counts


def split_detections(boxes, scores, classes, counts):
    """-> per image ``(boxes [D, 4], scores [D], classes [D])`` (one device-to-host read of the counts)."""
    n = counts.tolist()
    return [(boxes[i, :d], scores[i, :d], classes[i, :d]) for i, d in enumerate(n)]
