"""Generate tests/golden/target_cases.npz by EXECUTING the reference's own target-assignment and loss code
(run in the authoring container only; needs /root/reference, which is never read at test time).

    python tests/golden/gen_target_golden.py

The reference modules are imported BY FILE PATH with stub modules for what this container lacks (detectron2's
compiled extension, fvcore, yacs ...); nothing is copied or retyped.  What is executed, per SURVEY.md section 8(a):

  a11  RepPointsV2.point_targets            slender_det/modeling/meta_arch/reppoints/reppointsv2.py:370-428
  a12  RepPointsV2.bbox_targets             reppointsv2.py:430-484 (with the reference's Boxes / pairwise_iou)
  a13  compute_targets_for_locations        slender_det/modeling/meta_arch/fcos/utils.py:160-212 (+ get_sample_region)
       compute_topk_targets_for_locations   fcos/utils.py:215-292
  a18' compute_centerness_targets (slender) fcos/fcos_rpd_s1_topk.py:25-55   pow(c, min(w/h, h/w))
  a20  FCOSRepPoints.get_ground_truth       fcos/fcos_rpd_s1_topk.py:320-376 (its own top-5 targets :57-134 + Matcher)
       FCOSRepPoints.losses                 fcos/fcos_rpd_s1_topk.py:249-317

fvcore is a third-party dependency that is absent here (setup.py:106 pins only fvcore>=0.1.1): its two functions
on this path are stubbed with the published formulas -- torchvision.ops.sigmoid_focal_loss for
sigmoid_focal_loss_jit and the reference's own slender_det/layers/smooth_l1_loss_with_weight.py (which restates
fvcore's smooth_l1_loss) -- so those two terms are pinned on the stand-ins, everything else on the reference.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import REF, OUT, _load_by_path, _stub_module, _nonzero_tuple  # noqa: E402

MA = f"{REF}/slender_det/modeling/meta_arch"
LEVELS_SMALL = ((25, 42, 8), (13, 21, 16), (7, 11, 32), (4, 6, 64), (2, 3, 128))
LEVELS_FULL = ((100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128))


class _Registry:
    def register(self, obj=None):
        return obj if obj is not None else (lambda o: o)


def _any(*a, **k):
    raise RuntimeError("stubbed symbol called")


def load_target_reference():
    """-> namespace(boxes, matcher, utils, rpd, rpv2): the reference modules, imported by path."""
    from torchvision.ops import sigmoid_focal_loss
    boxes = _load_by_path("ref_boxes_t", f"{REF}/detectron2/detectron2/structures/boxes.py",
                          {"detectron2": _stub_module("detectron2"),
                           "detectron2.layers": _stub_module("detectron2.layers", nonzero_tuple=_nonzero_tuple)})
    matcher = _load_by_path("ref_matcher_t", f"{REF}/detectron2/detectron2/modeling/matcher.py")
    sl1 = _load_by_path("ref_sl1_t", f"{REF}/slender_det/layers/smooth_l1_loss_with_weight.py")
    iou = _load_by_path("ref_iou_t", f"{REF}/slender_det/layers/iou_loss.py")

    def smooth_l1_loss(input, target, beta, reduction="none"):
        return sl1.smooth_l1_loss_with_weight(input, target, None, beta, reduction=reduction)

    class Instances:  # attribute bag with the three fields the path reads
        def __init__(self, image_size, **kw):
            self.image_size = image_size
            self.__dict__.update(kw)

    cat = lambda ts, dim=0: ts[0] if len(ts) == 1 else torch.cat(ts, dim)
    stubs = {
        "fvcore": _stub_module("fvcore"),
        "fvcore.nn": _stub_module("fvcore.nn", sigmoid_focal_loss_jit=sigmoid_focal_loss, smooth_l1_loss=smooth_l1_loss),
        "detectron2": _stub_module("detectron2"),
        "detectron2.modeling": _stub_module("detectron2.modeling"),
        "detectron2.modeling.meta_arch": _stub_module("detectron2.modeling.meta_arch", META_ARCH_REGISTRY=_Registry()),
        "detectron2.modeling.backbone": _stub_module("detectron2.modeling.backbone", build_backbone=_any),
        "detectron2.modeling.matcher": _stub_module("detectron2.modeling.matcher", Matcher=matcher.Matcher),
        "detectron2.modeling.postprocessing": _stub_module("detectron2.modeling.postprocessing", detector_postprocess=_any),
        "detectron2.structures": _stub_module("detectron2.structures", Boxes=boxes.Boxes, pairwise_iou=boxes.pairwise_iou,
                                              ImageList=object, Instances=Instances),
        "detectron2.layers": _stub_module("detectron2.layers", cat=cat, ShapeSpec=object, batched_nms=_any,
                                          DeformConv=object, ModulatedDeformConv=object, nonzero_tuple=_nonzero_tuple),
        "detectron2.utils": _stub_module("detectron2.utils"),
        "detectron2.utils.logger": _stub_module("detectron2.utils.logger", log_first_n=lambda *a, **k: None),
        "slender_det": _stub_module("slender_det"),
        "slender_det.modeling": _stub_module("slender_det.modeling"),
        "slender_det.modeling.backbone": _stub_module("slender_det.modeling.backbone", build_backbone=_any),
        "slender_det.layers": _stub_module("slender_det.layers", Scale=object, iou_loss=iou.iou_loss, DFConv2d=object),
    }
    sys.modules.update(stubs)
    # the fcos files use `from .utils import ...`: give them a package whose __path__ is the reference directory,
    # WITHOUT running its __init__ (which would import every model file)
    pkg = types.ModuleType("ref_fcos_pkg")
    pkg.__path__ = [f"{MA}/fcos"]
    sys.modules["ref_fcos_pkg"] = pkg
    utils = importlib.import_module("ref_fcos_pkg.utils")
    rpd = importlib.import_module("ref_fcos_pkg.fcos_rpd_s1_topk")
    rpv2 = _load_by_path("ref_reppointsv2", f"{MA}/reppoints/reppointsv2.py")
    return types.SimpleNamespace(boxes=boxes, matcher=matcher, utils=utils, rpd=rpd, rpv2=rpv2, Instances=Instances)


# ---- seeded inputs ------------------------------------------------------------------------------------------------
def points_case(seed, M, levels):
    g = torch.Generator().manual_seed(seed)
    pts, strides = [], []
    for (h, w, s) in levels:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) * s, ys.reshape(-1) * s], 1))
        strides.append(torch.full((h * w,), float(s)))
    pts, strides = torch.cat(pts), torch.cat(strides)
    W, H = levels[0][1] * levels[0][2], levels[0][0] * levels[0][2]
    c = torch.rand(M, 2, generator=g) * torch.tensor([float(W), float(H)])
    wh = torch.exp(torch.rand(M, 2, generator=g) * 5.0 + 0.5)      # 1.6 .. 245 px: every level and both clamps
    gt = torch.cat([c - wh / 2, c + wh / 2], 1)
    gt[1] = gt[0]                                                  # two GTs at equal distance from one point
    return pts, strides, gt, torch.randint(0, 80, (M,), generator=g)


def bbox_case(seed, X, M):
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(X, 2, generator=g) * torch.tensor([640.0, 480.0])
    wh = torch.exp(torch.rand(X, 2, generator=g) * 3.5 + 1.5)
    cand = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)              # some coordinates negative: exercises the clamp
    c2 = torch.rand(M, 2, generator=g) * torch.tensor([640.0, 480.0])
    wh2 = torch.exp(torch.rand(M, 2, generator=g) * 3.5 + 2.0)
    gt = torch.cat([c2 - wh2 / 2, c2 + wh2 / 2], 1).clamp(min=0)
    gt[0] = torch.tensor([5000.0, 5000.0, 5100.0, 5100.0])        # a GT nothing overlaps: its maximum is 0
    return cand, gt, torch.randint(0, 80, (M,), generator=g)


def fcos_case(ref, seed, M, levels):
    g = torch.Generator().manual_seed(seed)
    shapes = [(h, w) for h, w, _ in levels]
    strides = [s for _, _, s in levels]
    locs = ref.utils.compute_locations(shapes, strides, "cpu")     # the reference's own location grid (:78-105)
    W, H = levels[0][1] * levels[0][2], levels[0][0] * levels[0][2]
    c = torch.rand(M, 2, generator=g) * torch.tensor([float(W), float(H)])
    wh = torch.exp(torch.rand(M, 2, generator=g) * 4.5 + 1.5)
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(min=0)
    boxes[3] = boxes[2]                                            # equal areas: the first index wins
    return locs, strides, boxes, torch.randint(0, 80, (M,), generator=g)


def soi_of(locs):
    INF = 100000000
    sizes = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, INF]]
    return torch.cat([l.new_tensor(sizes[i])[None].expand(len(l), -1) for i, l in enumerate(locs)], 0)


def main():
    torch.set_num_threads(4)
    ref = load_target_reference()
    out = {}
    put = lambda name, **kw: out.update({f"{name}/{k}": (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                         for k, v in kw.items()})
    v2 = types.SimpleNamespace(num_classes=80, point_base_scale=4)

    # ---- a11 point_targets ----
    for seed, M, lv, scale in ((0, 60, LEVELS_SMALL, 1.0), (1, 60, LEVELS_SMALL, 1.0), (3, 100, LEVELS_FULL, 1.0)):
        pts, strides, gt, labels = points_case(seed, M, lv)
        b, l = ref.rpv2.RepPointsV2.point_targets(v2, pts, strides, gt, labels)
        assert (l != 80).sum() > 10
        put(f"pt{seed}", points=pts, strides=strides, gt=gt, labels=labels, boxes=b, assigned=l)

    # ---- a12 bbox_targets ----
    for seed, X, M, gmm in ((0, 3000, 23, True), (1, 3000, 23, False), (7, 22400, 100, True)):
        cand, gt, labels = bbox_case(seed, X, M)
        c_in = cand.clone()
        b, l = ref.rpv2.RepPointsV2.bbox_targets(v2, c_in, ref.boxes.Boxes(gt), labels, gt_max_matching=gmm)
        put(f"bb{seed}", cand=cand, cand_clamped=c_in, gt=gt, labels=labels, gmm=np.array(gmm), boxes=b, assigned=l)

    # ---- a13 FCOS location targets (+ top-k variant, sqrt centerness) ----
    for seed, M, lv, radius in ((0, 40, LEVELS_SMALL, 0.0), (1, 40, LEVELS_SMALL, 1.5), (2, 40, LEVELS_SMALL, 1.0),
                                (5, 100, LEVELS_FULL, 1.5)):
        locs, strides, boxes, classes = fcos_case(ref, seed, M, lv)
        soi = soi_of(locs)
        inst = [ref.Instances((0, 0), gt_boxes=ref.boxes.Boxes(boxes), gt_classes=classes)]
        cls, reg = ref.utils.compute_targets_for_locations(locs, inst, soi, strides, radius, 80)
        assert (cls != 80).sum() > 20
        for norm in (False, True):
            c2, r2, tk = ref.utils.compute_topk_targets_for_locations(locs, inst, soi, strides, radius, 80,
                                                                      norm_reg_targets=norm, topk=5)
            assert torch.equal(c2, cls)
            put(f"fc{seed}", **{f"topk_reg{int(norm)}": r2[0], f"topk_mask{int(norm)}": tk[0]})
        put(f"fc{seed}", boxes=boxes, classes=classes, radius=np.array(radius), levels=np.array(lv),
            out_classes=cls[0], out_reg=reg[0])
    # the "first GT centred at x == 0" shortcut of get_sample_region (:122-123)
    locs, strides, boxes, classes = fcos_case(ref, 6, 40, LEVELS_SMALL)
    boxes[0] = torch.tensor([-20.0, 10.0, 20.0, 60.0])
    inst = [ref.Instances((0, 0), gt_boxes=ref.boxes.Boxes(boxes), gt_classes=classes)]
    cls, reg = ref.utils.compute_targets_for_locations(locs, inst, soi_of(locs), strides, 1.5, 80)
    assert (cls == 80).all()
    put("fc6", boxes=boxes, classes=classes, radius=np.array(1.5), levels=np.array(LEVELS_SMALL), out_classes=cls[0],
        out_reg=reg[0])

    # ---- slender centerness (fcos_rpd_s1_topk.py:25-55) and the FCOS one (fcos/utils.py:295-300) ----
    g = torch.Generator().manual_seed(31)
    ltrb = torch.exp(torch.rand(4099, 4, generator=g) * 6.0 - 1.0)
    put("ctr", ltrb=ltrb, slender=ref.rpd.compute_centerness_targets(ltrb), fcos=ref.utils.compute_centerness_targets(ltrb))

    # ---- a20 FCOSRepPoints.get_ground_truth + losses ----
    me = types.SimpleNamespace(num_classes=80, fpn_strides=[8, 16, 32, 64, 128], center_sampling_radius=0.0,
                               focal_loss_alpha=0.25, focal_loss_gamma=2.0, iou_loss_type="iou",
                               bbox_matcher=ref.matcher.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=True))
    for tag, lv, M, radius in (("rpd_small", LEVELS_SMALL, 30, 0.0), ("rpd_full", LEVELS_FULL, 50, 0.0),
                               ("rpd_cs", LEVELS_SMALL, 30, 1.5)):
        me.center_sampling_radius = radius
        g = torch.Generator().manual_seed(21)
        sizes = [(lv[0][0] * 8, lv[0][1] * 8 - 11), (lv[0][0] * 8 - 96, lv[0][1] * 8 - 128)]
        cases = [fcos_case(ref, s, M, lv) for s in (12, 13)]
        locs, strides = cases[0][0], cases[0][1]
        centers = torch.cat(locs)
        X = centers.shape[0]
        init = []
        for _ in cases:
            wh = torch.exp(torch.rand(X, 2, generator=g) * 3.5 + 2.0)
            init.append(torch.cat([centers - wh / 2, centers + wh / 2], 1))
        inst = [ref.Instances(sz, gt_boxes=ref.boxes.Boxes(c[2]), gt_classes=c[3]) for c, sz in zip(cases, sizes)]
        ic, ir, rc, rr, tk = ref.rpd.FCOSRepPoints.get_ground_truth(me, locs, init, inst)
        assert (rc == -1).any() and (rc == 80).any() and ((rc >= 0) & (rc < 80)).any() and tk.sum() > 20
        put(tag, levels=np.array(lv), radius=np.array(radius), sizes=np.array(sizes),
            boxes0=cases[0][2], classes0=cases[0][3], boxes1=cases[1][2], classes1=cases[1][3],
            init0=init[0], init1=init[1], init_classes=ic, init_reg=ir, refine_classes=rc, refine_reg=rr, topk=tk)
        if tag != "rpd_small":
            continue
        # losses on those targets: level-shaped predictions as the head emits them (N, K|4|1, H, W)
        N = 2
        mk = lambda ch, scale, shift=0.0: [(torch.randn(N, ch, h, w, generator=g) * scale + shift).requires_grad_()
                                           for h, w, _ in lv]
        box_cls = mk(80, 2.0, -4.6)
        # positive ltrb predictions near the targets keep -log(iou) finite, as in training
        box_init = [t.detach().abs().add(1.0).requires_grad_() for t in mk(4, 30.0)]
        box_ref = mk(4, 30.0)
        ctr = mk(1, 1.5)
        strides_t = torch.cat([torch.full((h * w,), float(s)) for h, w, s in lv])
        losses = ref.rpd.FCOSRepPoints.losses(me, ic, ir, rc, rr, box_cls, box_init, box_ref, ctr, strides_t, tk)
        total = sum(losses.values())
        total.backward()
        flat = lambda ts, k: torch.cat([t.permute(0, 2, 3, 1).reshape(N, -1, k) for t in ts], 1).reshape(-1, k)
        put(tag + "_loss", logits=flat(box_cls, 80), box_init=flat(box_init, 4), box_ref=flat(box_ref, 4),
            ctr=flat(ctr, 1).reshape(-1), strides=strides_t.repeat(N),
            g_logits=flat([t.grad for t in box_cls], 80), g_box_init=flat([t.grad for t in box_init], 4),
            g_box_ref=flat([t.grad for t in box_ref], 4), g_ctr=flat([t.grad for t in ctr], 1).reshape(-1),
            **{k: v for k, v in losses.items()})
    np.savez_compressed(os.path.join(OUT, "target_cases.npz"), **out)
    print("target_cases.npz:", len(out), "arrays,", os.path.getsize(os.path.join(OUT, "target_cases.npz")) >> 10, "KiB")


if __name__ == "__main__":
    main()
