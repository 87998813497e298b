"""Label assignment behind the reference's matcher API, fused into libslender_b200 kernels.

  pairwise_iou   detectron2/structures/boxes.py:316-348 (accepts Boxes-like objects or [N,4] tensors)
  Matcher        detectron2/modeling/matcher.py:8-126
  TopKMatcher    slender_det/modeling/matchers/topk_matcher.py:7-86
  iou_assign     fused pairwise_iou + (TopK)Matcher straight from boxes: the [M,X] IoU matrix is
                 never written to HBM (it is ~9 MB per image in the reference).

Outputs are ``matches`` int64 [X] and ``match_labels`` int8 [X], bit-identical to the reference on
tie-free inputs; TopK ties resolve to the lowest anchor index (see oracle/assign.py).  Unlike the
reference there is no host-synchronising ``assert torch.all(q >= 0)``.
"""
import ctypes
from typing import List

import torch

from . import _lib


def _tensor_of(boxes):
    return boxes.tensor if hasattr(boxes, "tensor") else boxes


def pairwise_iou(boxes1, boxes2):
    b1 = _tensor_of(boxes1).detach().float().contiguous()
    b2 = _tensor_of(boxes2).detach().float().contiguous()
    if not b1.is_cuda:
        raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(_lib.lib().sdb_pairwise_iou(_lib.ptr(b1), _lib.ptr(b2), b1.shape[0], b2.shape[0],
                                               _lib.ptr(out), _lib.stream_ptr(b1.device)))
    return out


def _cfg_arrays(thresholds, labels):
    th = (ctypes.c_float * len(thresholds))(*[float(t) for t in thresholds])
    lb = (ctypes.c_int8 * len(labels))(*[int(l) for l in labels])
    return th, lb


class _MatcherBase(object):
    def __init__(self, thresholds: List[float], labels: List[int]):
        thresholds = thresholds[:]
        assert thresholds[0] > 0
        self._user_thresholds = thresholds[:]
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all([low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:])])
        assert all([l in [-1, 0, 1] for l in labels])
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels

    def _run(self, q, topk, alq):
        assert q.dim() == 2
        if not q.is_cuda:
            raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
        q = q.detach().float().contiguous()
        M, X = q.shape
        matches = torch.empty((X,), dtype=torch.int64, device=q.device)
        mlabels = torch.empty((X,), dtype=torch.int8, device=q.device)
        th, lb = _cfg_arrays(self._user_thresholds, self.labels)
        wsb = int(_lib.lib().sdb_assign_workspace_bytes(M, X, topk))
        ws = torch.empty(wsb, dtype=torch.uint8, device=q.device) if wsb else None
        with torch.cuda.device(q.device):
            _lib.check(_lib.lib().sdb_match_quality_assign(_lib.ptr(q), M, X, th, lb, len(self._user_thresholds),
                                                           topk, int(alq), _lib.ptr(matches), _lib.ptr(mlabels),
                                                           _lib.ptr(ws), wsb, _lib.stream_ptr(q.device)))
        return matches, mlabels

    def _run_boxes(self, gt, anchors, topk, alq, return_iou=False):
        gt = _tensor_of(gt).detach().float().contiguous()
        an = _tensor_of(anchors).detach().float().contiguous()
        if not an.is_cuda:
            raise RuntimeError("slender_b200: CUDA tensors only (no CPU fallback)")
        M, X = gt.shape[0], an.shape[0]
        matches = torch.empty((X,), dtype=torch.int64, device=an.device)
        mlabels = torch.empty((X,), dtype=torch.int8, device=an.device)
        iou = torch.empty((M, X), dtype=torch.float32, device=an.device) if return_iou else None
        th, lb = _cfg_arrays(self._user_thresholds, self.labels)
        wsb = int(_lib.lib().sdb_assign_workspace_bytes(M, X, topk))
        ws = torch.empty(wsb, dtype=torch.uint8, device=an.device) if wsb else None
        with torch.cuda.device(an.device):
            _lib.check(_lib.lib().sdb_iou_assign(_lib.ptr(gt), _lib.ptr(an), M, X, th, lb,
                                                 len(self._user_thresholds), topk, int(alq), _lib.ptr(matches),
                                                 _lib.ptr(mlabels), _lib.ptr(iou), _lib.ptr(ws), wsb,
                                                 _lib.stream_ptr(an.device)))
        return (matches, mlabels, iou) if return_iou else (matches, mlabels)


class Matcher(_MatcherBase):
    """detectron2.modeling.matcher.Matcher: argmax-GT per prediction, threshold labels, optional
    low-quality matches."""

    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False):
        super().__init__(thresholds, labels)
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix):
        return self._run(match_quality_matrix, 0, self.allow_low_quality_matches)

    def from_boxes(self, gt_boxes, anchors, return_iou=False):
        """Fused pairwise_iou(gt, anchors) + __call__; the IoU matrix stays on chip."""
        return self._run_boxes(gt_boxes, anchors, 0, self.allow_low_quality_matches, return_iou)


class TopKMatcher(_MatcherBase):
    """slender_det TopKMatcher: Matcher labels, then every GT's top-k anchors are forced positive."""

    def __init__(self, thresholds: List[float], labels: List[int], topk: int = 9):
        super().__init__(thresholds, labels)
        self.topk = topk

    def __call__(self, match_quality_matrix):
        return self._run(match_quality_matrix, self.topk, False)

    def from_boxes(self, gt_boxes, anchors, return_iou=False):
        return self._run_boxes(gt_boxes, anchors, self.topk, False, return_iou)
