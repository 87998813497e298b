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


# ---------------------------------------------------------------------------------------------------------------------
# Target assignment: the numpy oracle against tests/golden/target_cases.npz, which
# tests/golden/gen_target_golden.py produced by EXECUTING the reference's own functions (imported by file path):
# RepPointsV2.point_targets / bbox_targets (reppointsv2.py:370-484), compute_targets_for_locations /
# compute_topk_targets_for_locations (fcos/utils.py:160-292), FCOSRepPoints.get_ground_truth
# (fcos_rpd_s1_topk.py:320-376) and its module-local top-5 targets and pow-centerness (:25-134).
# ---------------------------------------------------------------------------------------------------------------------
from target_cases import locations_of, soi_of, strides_of, num_points_of  # noqa: E402


@pytest.mark.parametrize("name", ["pt0", "pt1", "pt3"])
def test_point_targets_oracle_vs_reference(target_cases, name):
    c = target_cases[name]
    ob, ol_ = oa.point_targets(c["points"], c["strides"], c["gt"], c["labels"], 80)
    assert np.array_equal(ol_, c["assigned"]) and np.array_equal(ob, c["boxes"])
    assert (c["assigned"] != 80).sum() > 10


@pytest.mark.parametrize("name", ["bb0", "bb1", "bb7"])
def test_bbox_targets_oracle_vs_reference(target_cases, name):
    c = target_cases[name]
    cand = c["cand"].copy()
    ob, ol_ = oa.bbox_targets(cand, c["gt"], c["labels"], 80, gt_max_matching=bool(c["gmm"]))
    assert np.array_equal(cand, c["cand_clamped"])                # same in-place clamp (reppointsv2.py:452-455)
    assert np.array_equal(ol_, c["assigned"]) and np.array_equal(ob, c["boxes"])


@pytest.mark.parametrize("name", ["fc0", "fc1", "fc2", "fc5", "fc6"])
def test_fcos_location_targets_oracle_vs_reference(target_cases, name):
    c = target_cases[name]
    lv = c["levels"]
    loc = np.concatenate(locations_of(lv))
    oc, orr, idx = oa.fcos_location_targets(loc, soi_of(lv), c["boxes"], c["classes"], num_points_of(lv), strides_of(lv),
                                            float(c["radius"]), 80, return_index=True)
    assert np.array_equal(oc, c["out_classes"]) and np.array_equal(orr, c["out_reg"])
    if "topk_mask0" in c:
        ot = oa.fcos_topk_locations(oc, orr, idx, 80, topk=5)
        assert np.array_equal(ot, c["topk_mask0"]) and np.array_equal(ot, c["topk_mask1"]) and ot.sum() > 20
        assert np.array_equal(orr, c["topk_reg0"])
        norm = np.concatenate([np.full(n, s, np.float32) for n, s in zip(num_points_of(lv), strides_of(lv))])
        assert np.array_equal(orr / norm[:, None], c["topk_reg1"])


def test_centerness_oracles_vs_reference(target_cases):
    """fcos/utils.py:295-300 (sqrt) and the FCOSRepPoints module's own pow(c, min(w/h, h/w)) (fcos_rpd_s1_topk.py:25-55)."""
    from oracle import losses as ol
    c = target_cases["ctr"]
    assert np.allclose(ol.centerness_targets(c["ltrb"]).float().numpy(), c["fcos"], rtol=2e-7, atol=0)
    got = ol.slender_centerness_targets(c["ltrb"]).float().numpy()
    assert np.allclose(got, c["slender"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("name", ["rpd_small", "rpd_full", "rpd_cs"])
def test_fcos_rpd_ground_truth_oracle_vs_reference(target_cases, name):
    c = target_cases[name]
    lv = c["levels"]
    loc = np.concatenate(locations_of(lv))
    for i in range(2):
        oc, orr, idx = oa.fcos_location_targets(loc, soi_of(lv), c["boxes%d" % i], c["classes%d" % i], num_points_of(lv),
                                                strides_of(lv), float(c["radius"]), 80, return_index=True)
        assert np.array_equal(oc, c["init_classes"][i]) and np.array_equal(orr, c["init_reg"][i])
        assert np.array_equal(oa.fcos_topk_locations(oc, orr, idx, 80, topk=5, slender=True), c["topk"][i])
        c2, r2 = oa.fcos_rpd_refine_targets(loc, c["init%d" % i], c["boxes%d" % i], c["classes%d" % i],
                                            tuple(int(v) for v in c["sizes"][i]), 80)
        assert np.array_equal(c2, c["refine_classes"][i]) and np.array_equal(r2, c["refine_reg"][i])
