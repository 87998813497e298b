"""Generate tests/golden/postproc_cases.npz by EXECUTING the reference's RepPointsV2.inference_single_image
(/root/reference/slender_det/modeling/meta_arch/reppoints/reppointsv2.py:533-603, with pts_to_bbox :328-366) by file path
on seeded inputs (authoring container only; /root/reference is never read at test time).

    python tests/golden/gen_postproc_golden.py

detectron2.layers.batched_nms (detectron2/layers/nms.py:10-29) forwards to torchvision.ops.boxes.batched_nms for fewer
than 40 000 boxes; the compiled detectron2 extension is absent here, so the stub binds torchvision's function directly.
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT  # noqa: E402
import gen_target_golden as G  # noqa: E402


def make_case(seed, levels, K, npts=9, hot=0.02):
    """levels: (H, W, stride).  Logits ~ prior 0.01 with a fraction `hot` of confident cells; refined point sets around
    the centre, a few stretching outside the image."""
    g = torch.Generator().manual_seed(seed)
    cls, pts, ctr, strides = [], [], [], []
    for (h, w, s) in levels:
        n = h * w
        logit = torch.randn(n, K, generator=g) * 1.5 - 4.6
        m = torch.rand(n, K, generator=g) < hot
        logit[m] = torch.randn(int(m.sum()), generator=g) * 2.0 + 0.5
        cls.append(logit)
        pts.append(torch.randn(n, 2 * npts, generator=g) * 2.5)
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32) * s, torch.arange(w, dtype=torch.float32) * s, indexing="ij")
        ctr.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], 1))
        strides.append(torch.full((n,), float(s)))
    return cls, pts, ctr, strides


def main():
    ref = G.load_target_reference()
    rp = ref.rpv2
    rp.batched_nms = torchvision.ops.boxes.batched_nms          # what detectron2.layers.batched_nms calls (< 40 000 boxes)
    out = {}
    cases = [("small", 1, ((25, 42, 8), (13, 21, 16), (7, 11, 32)), 20, dict(topk=1000, thr=0.05, nms=0.5, det=100), "minmax"),
             ("topk_binds", 2, ((25, 42, 8), (13, 21, 16), (7, 11, 32)), 20, dict(topk=60, thr=0.05, nms=0.5, det=40), "minmax"),
             ("partial", 3, ((13, 21, 16), (7, 11, 32)), 8, dict(topk=1000, thr=0.3, nms=0.6, det=100), "partial_minmax"),
             ("moment", 4, ((13, 21, 16), (7, 11, 32)), 8, dict(topk=1000, thr=0.05, nms=0.5, det=100), "moment")]
    for name, seed, levels, K, cfg, tm in cases:
        cls, pts, ctr, strides = make_case(seed, levels, K, hot=0.05 if name == "topk_binds" else 0.02)
        me = types.SimpleNamespace(num_classes=K, topk_candidates=cfg["topk"], score_threshold=cfg["thr"],
                                   nms_threshold=cfg["nms"], max_detections_per_image=cfg["det"], transform_method=tm,
                                   moment_transfer=torch.tensor([0.3, -0.2]), moment_mul=0.01)
        me.pts_to_bbox = types.MethodType(rp.RepPointsV2.pts_to_bbox, me)
        H, W, s0 = levels[0]
        image_size = (H * s0 - 9, W * s0 - 21)
        res = rp.RepPointsV2.inference_single_image(me, [c.clone() for c in cls], pts, strides, ctr, image_size)
        scores = res.scores.numpy()
        assert len(scores) > 10 and (np.diff(scores) <= 0).all()
        assert len(np.unique(scores)) == len(scores), "tie-free case expected"
        for l in range(len(levels)):
            out[f"{name}/cls{l}"], out[f"{name}/pts{l}"] = cls[l].numpy(), pts[l].numpy()
        out[f"{name}/levels"] = np.array(levels)
        out[f"{name}/cfg"] = np.array([K, cfg["topk"], cfg["det"]])
        out[f"{name}/fcfg"] = np.array([cfg["thr"], cfg["nms"], 0.3, -0.2], np.float32)
        out[f"{name}/transform"] = np.array(["minmax", "partial_minmax", "moment"].index(tm))
        out[f"{name}/image_size"] = np.array(image_size)
        out[f"{name}/boxes"] = res.pred_boxes.tensor.numpy()
        out[f"{name}/scores"] = scores
        out[f"{name}/classes"] = res.pred_classes.numpy()
    np.savez_compressed(os.path.join(OUT, "postproc_cases.npz"), **out)
    print("postproc_cases.npz:", len(out), "arrays,", os.path.getsize(os.path.join(OUT, "postproc_cases.npz")) >> 10, "KiB")


if __name__ == "__main__":
    main()
