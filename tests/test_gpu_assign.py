"""GPU parity: fused IoU + label assignment, bit-exact against goldens from the reference's Python."""
import numpy as np
import pytest
import torch

import slenderobjdet_b200 as sdb
from oracle import assign as oa

pytestmark = pytest.mark.gpu


def _d(a):
    return torch.as_tensor(a).cuda()


def test_known_answers(assign_cases):
    c = assign_cases["ka_iou"]
    assert np.allclose(sdb.pairwise_iou(_d(c["boxes1"]), _d(c["boxes2"])).cpu().numpy(), c["expected"])
    c = assign_cases["ka_matcher"]
    m, l = sdb.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(_d(c["q"]))
    assert m.dtype == torch.int64 and l.dtype == torch.int8
    assert np.array_equal(m.cpu().numpy(), c["matches"]) and np.array_equal(l.cpu().numpy(), c["labels"])


@pytest.mark.parametrize("name", ["small", "mid", "slender"])
def test_iou_bit_exact(assign_cases, name):
    c = assign_cases[name]
    iou = sdb.pairwise_iou(_d(c["gt"]), _d(c["anchors"])).cpu().numpy()
    assert np.array_equal(iou.view(np.uint32), c["iou"].view(np.uint32))
    iou_t = sdb.pairwise_iou(_d(c["anchors"]), _d(c["gt"])).cpu().numpy()  # [X,M] as bbox_targets calls it
    assert np.array_equal(iou_t.view(np.uint32), c["iou"].T.view(np.uint32))


@pytest.mark.parametrize("name", ["small", "mid", "slender"])
@pytest.mark.parametrize("k", [1, 9, 10])
def test_topk_matcher_bit_exact(assign_cases, name, k):
    c = assign_cases[name]
    assert bool(c[f"topk{k}_tiefree"])
    tm = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=k)
    m, l = tm(_d(c["iou"]))
    assert np.array_equal(m.cpu().numpy(), c[f"topk{k}_matches"])
    assert np.array_equal(l.cpu().numpy(), c[f"topk{k}_labels"])
    m2, l2, iou = tm.from_boxes(_d(c["gt"]), _d(c["anchors"]), return_iou=True)  # fused, no IoU matrix needed
    assert np.array_equal(m2.cpu().numpy(), c[f"topk{k}_matches"])
    assert np.array_equal(l2.cpu().numpy(), c[f"topk{k}_labels"])
    assert np.array_equal(iou.cpu().numpy().view(np.uint32), c["iou"].view(np.uint32))


@pytest.mark.parametrize("name", ["small", "mid", "slender"])
@pytest.mark.parametrize("alq", [0, 1])
def test_matcher_bit_exact(assign_cases, name, alq):
    c = assign_cases[name]
    mt = sdb.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=bool(alq))
    for m, l in (mt(_d(c["iou"])), mt.from_boxes(_d(c["gt"]), _d(c["anchors"]))):
        assert np.array_equal(m.cpu().numpy(), c[f"matcher{alq}_matches"])
        assert np.array_equal(l.cpu().numpy(), c[f"matcher{alq}_labels"])


def test_empty_gt_and_k_too_large(assign_cases):
    c = assign_cases["empty"]
    m, l = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=9)(torch.zeros(0, 11, device="cuda"))
    assert np.array_equal(m.cpu().numpy(), c["matches"]) and np.array_equal(l.cpu().numpy(), c["labels"])
    with pytest.raises(RuntimeError):  # torch.topk: k out of range
        sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=9)(torch.rand(2, 5, device="cuda"))


def test_ties_follow_canonical_rule_and_full_size():
    """BASELINE config 4: ~22 400 points x 100 GT, k in {9,10}; ties -> lowest anchor index."""
    g = torch.Generator().manual_seed(11)
    X, M = 22400, 100
    ctr = torch.rand(X, 2, generator=g) * torch.tensor([1333.0, 800.0])
    an = torch.cat([ctr - 16, ctr + 16], 1)
    c2 = torch.rand(M, 2, generator=g) * torch.tensor([1333.0, 800.0])
    wh = torch.exp(torch.rand(M, 2, generator=g) * 4 + 2)
    gt = torch.cat([c2 - wh / 2, c2 + wh / 2], 1)
    gt[0] = torch.tensor([-500.0, -500.0, -400.0, -400.0])  # overlaps nothing: an all-zero (all-tie) row
    q = oa.pairwise_iou(gt.numpy(), an.numpy())
    for k in (9, 10):
        mo, lo = oa.topk_matcher(q, [0.3, 0.7], [0, -1, 1], topk=k)
        m, l = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=k).from_boxes(gt.cuda(), an.cuda())
        assert np.array_equal(m.cpu().numpy(), mo) and np.array_equal(l.cpu().numpy(), lo)
        assert (lo[:k] == 1).all()  # the all-tie row picks anchors 0..k-1
        # same through the matcher's own signature (quality matrix), and on a size the 8 scan splits do not divide
        m2, l2 = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=k)(torch.from_numpy(q).cuda())
        assert np.array_equal(m2.cpu().numpy(), mo) and np.array_equal(l2.cpu().numpy(), lo)
        q3 = q[:, :22397]
        mo3, lo3 = oa.topk_matcher(q3, [0.3, 0.7], [0, -1, 1], topk=k)
        m3, l3 = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=k)(torch.from_numpy(np.ascontiguousarray(q3)).cuda())
        assert np.array_equal(m3.cpu().numpy(), mo3) and np.array_equal(l3.cpu().numpy(), lo3)
    mo, lo = oa.matcher(q, [0.4, 0.5], [0, -1, 1], True)
    m, l = sdb.Matcher([0.4, 0.5], [0, -1, 1], True).from_boxes(gt.cuda(), an.cuda())
    assert np.array_equal(m.cpu().numpy(), mo) and np.array_equal(l.cpu().numpy(), lo)


# ---------------------------------------------------------------------------------------------------------------------
# Target assignment on the GPU against the REFERENCE's own outputs (tests/golden/target_cases.npz, produced by
# tests/golden/gen_target_golden.py executing reppointsv2.py / fcos/utils.py / fcos_rpd_s1_topk.py by file path).
# ---------------------------------------------------------------------------------------------------------------------
from target_cases import locations_of, soi_of, strides_of  # noqa: E402


@pytest.mark.parametrize("name", ["bb0", "bb1", "bb7"])
def test_bbox_targets_bit_exact_vs_reference(target_cases, name):
    """RepPointsV2.bbox_targets (reppointsv2.py:430-484) on the fused IoU/matcher kernel: labels and boxes
    bit-identical to the reference's output, candidates clamped in place, no IoU matrix."""
    from slenderobjdet_b200.targets import bbox_targets
    c = target_cases[name]
    cd = _d(c["cand"].copy())
    b, l = bbox_targets(cd, _d(c["gt"]), _d(c["labels"]), 80, gt_max_matching=bool(c["gmm"]))
    assert np.array_equal(cd.cpu().numpy(), c["cand_clamped"])
    assert l.dtype == torch.int64 and np.array_equal(l.cpu().numpy(), c["assigned"])
    assert np.array_equal(b.cpu().numpy(), c["boxes"])
    with pytest.raises(ValueError):
        bbox_targets(torch.zeros(0, 4, device="cuda"), _d(c["gt"]), _d(c["labels"]), 80)


@pytest.mark.parametrize("name", ["pt0", "pt1", "pt3"])
def test_point_targets_bit_exact_vs_reference(target_cases, name):
    """RepPointsV2.point_targets (reppointsv2.py:370-428): three launches instead of the host loop over GTs; boxes and
    labels identical to the reference's, incl. two GTs at equal distance from one point (the first GT keeps it) and
    the full P3-P7 point set of an 800x1344 image (pt3)."""
    from slenderobjdet_b200.targets import point_targets
    c = target_cases[name]
    b, l = point_targets(_d(c["points"]), _d(c["strides"]), _d(c["gt"]), _d(c["labels"]), 80)
    assert l.dtype == torch.int64 and np.array_equal(l.cpu().numpy(), c["assigned"])
    assert np.array_equal(b.cpu().numpy(), c["boxes"])
    with pytest.raises(ValueError):
        point_targets(_d(c["points"][:0]), _d(c["strides"][:0]), _d(c["gt"]), _d(c["labels"]), 80)


def _fcos_inputs(c):
    lv = c["levels"]
    return [_d(l) for l in locations_of(lv)], _d(soi_of(lv)), strides_of(lv)


@pytest.mark.parametrize("names", [("fc0",), ("fc1",), ("fc2",), ("fc5",), ("fc6",), ("fc1", "fc1")])
def test_fcos_location_targets_bit_exact_vs_reference(target_cases, names):
    """compute_targets_for_locations / compute_topk_targets_for_locations (fcos/utils.py:160-292), one batched call:
    classes, ltrb targets (optionally stride-normalised) and the per-GT top-5-by-centerness mask identical to the
    reference's output; with and without centre sampling; equal-area ties; the P3-P7 grid (fc5); get_sample_region's
    "first GT centred at x == 0" shortcut (fc6); images with different GT counts in one batch share the padded call."""
    from slenderobjdet_b200.targets import compute_targets_for_locations, compute_topk_targets_for_locations
    cs = [target_cases[n] for n in names]
    locs, soi, strides = _fcos_inputs(cs[0])
    radius = float(cs[0]["radius"])
    targets = [(_d(c["boxes"]), _d(c["classes"])) for c in cs]
    if len(cs) > 1:   # ragged batch: the second image keeps only its first 17 GTs; checked against the oracle
        targets[1] = (targets[1][0][:17].contiguous(), targets[1][1][:17].contiguous())
    cls, reg = compute_targets_for_locations(locs, targets, soi, strides, radius, 80)
    assert cls.shape == (len(cs), soi.shape[0]) and reg.shape == (len(cs), soi.shape[0], 4) and cls.dtype == torch.int64
    assert np.array_equal(cls[0].cpu().numpy(), cs[0]["out_classes"]) and np.array_equal(reg[0].cpu().numpy(), cs[0]["out_reg"])
    if len(cs) > 1:
        lv = cs[0]["levels"]
        oc, orr = oa.fcos_location_targets(np.concatenate(locations_of(lv)), soi_of(lv), cs[1]["boxes"][:17],
                                           cs[1]["classes"][:17], [len(l) for l in locs], strides, radius, 80)
        assert np.array_equal(cls[1].cpu().numpy(), oc) and np.array_equal(reg[1].cpu().numpy(), orr)
    if "topk_mask0" in cs[0]:
        for norm in (0, 1):
            c2, r2, tk = compute_topk_targets_for_locations(locs, targets, soi, strides, radius, 80,
                                                            norm_reg_targets=bool(norm), topk=5)
            assert tk.dtype == torch.bool and np.array_equal(tk[0].cpu().numpy(), cs[0]["topk_mask%d" % norm])
            assert np.array_equal(c2[0].cpu().numpy(), cs[0]["out_classes"])
            assert np.array_equal(r2[0].cpu().numpy(), cs[0]["topk_reg%d" % norm])


@pytest.mark.parametrize("name", ["rpd_small", "rpd_full", "rpd_cs"])
def test_fcos_rpd_get_ground_truth_bit_exact_vs_reference(target_cases, name):
    """FCOSRepPoints.get_ground_truth (fcos_rpd_s1_topk.py:320-376), both stages, against the reference's output:
    stage-1 classes / ltrb / top-5 mask (ranked by the module's own pow-form centerness) and stage-2 IoU-matched
    classes (incl. -1 outside the image, num_classes for unmatched) / ltrb."""
    from slenderobjdet_b200.targets import fcos_rpd_get_ground_truth
    c = target_cases[name]
    locs, _, strides = _fcos_inputs(c)
    gts = [(_d(c["boxes%d" % i]), _d(c["classes%d" % i]), tuple(int(v) for v in c["sizes"][i])) for i in range(2)]
    ic, ir, rc, rr, tk = fcos_rpd_get_ground_truth(locs, [_d(c["init0"]), _d(c["init1"])], gts, strides, float(c["radius"]), 80)
    assert np.array_equal(ic.cpu().numpy(), c["init_classes"]) and np.array_equal(ir.cpu().numpy(), c["init_reg"])
    assert np.array_equal(tk.cpu().numpy(), c["topk"])
    assert np.array_equal(rc.cpu().numpy(), c["refine_classes"]) and np.array_equal(rr.cpu().numpy(), c["refine_reg"])
