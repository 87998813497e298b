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


@pytest.mark.parametrize("gmm", [True, False])
@pytest.mark.parametrize("size", [(3000, 23), (22400, 100)])
def test_bbox_targets_bit_exact(size, gmm):
    """RepPointsV2.bbox_targets (reppointsv2.py:430-484) on the fused IoU/matcher kernel: labels and boxes
    bit-identical to the oracle, candidates clamped in place, no IoU matrix."""
    from test_oracle_assign import _bbox_case
    from slenderobjdet_b200.targets import bbox_targets
    cand, gt, labels = _bbox_case(7, *size)
    cd = cand.clone().cuda()
    b, l = bbox_targets(cd, gt.cuda(), labels.cuda(), 80, gt_max_matching=gmm)
    cn = cand.clone().numpy()
    ob, ol_ = oa.bbox_targets(cn, gt.numpy(), labels.numpy(), 80, gt_max_matching=gmm)
    assert np.array_equal(cd.cpu().numpy(), cn)
    assert l.dtype == torch.int64 and np.array_equal(l.cpu().numpy(), ol_)
    assert np.array_equal(b.cpu().numpy(), ob)
    with pytest.raises(ValueError):
        bbox_targets(torch.zeros(0, 4, device="cuda"), gt.cuda(), labels.cuda(), 80)


@pytest.mark.parametrize("seed", [0, 3])
def test_point_targets_bit_exact(seed):
    """RepPointsV2.point_targets (reppointsv2.py:370-428): three launches instead of the host loop over GTs;
    boxes and labels identical to the oracle, incl. two GTs at equal distance from one point (first GT keeps it)
    and the full P3-P7 point set of an 800x1344 image."""
    from test_oracle_assign import _points_case
    from slenderobjdet_b200.targets import point_targets
    lv = ((100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128)) if seed else None
    pts, strides, gt, labels = _points_case(seed, 100, lv) if lv else _points_case(seed)
    if lv:
        gt = gt * 4.0
    b, l = point_targets(pts.cuda(), strides.cuda(), gt.cuda(), labels.cuda(), 80)
    ob, ol_ = oa.point_targets(pts.numpy(), strides.numpy(), gt.numpy(), labels.numpy(), 80)
    assert l.dtype == torch.int64 and np.array_equal(l.cpu().numpy(), ol_)
    assert np.array_equal(b.cpu().numpy(), ob)
    with pytest.raises(ValueError):
        point_targets(pts[:0].cuda(), strides[:0].cuda(), gt.cuda(), labels.cuda(), 80)


@pytest.mark.parametrize("radius", [0.0, 1.5])
@pytest.mark.parametrize("full", [False, True])
def test_fcos_location_targets_bit_exact(radius, full):
    """compute_targets_for_locations (fcos/utils.py:160-212), one fused kernel per image: classes and ltrb targets
    identical to the oracle, with and without centre sampling, two images, equal-area ties, the P3-P7 grid."""
    from test_oracle_assign import _fcos_case
    from slenderobjdet_b200.targets import compute_targets_for_locations
    lv = ((100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128))
    cases = [_fcos_case(s, 100, lv) if full else _fcos_case(s) for s in (4, 5)]
    locs, soi, _, _, strides = cases[0]
    targets = [(c[2].cuda(), c[3].cuda()) for c in cases]
    cls, reg = compute_targets_for_locations([l.cuda() for l in locs], targets, soi.cuda(), strides, radius, 80)
    assert cls.shape == (2, soi.shape[0]) and reg.shape == (2, soi.shape[0], 4) and cls.dtype == torch.int64
    for i, c in enumerate(cases):
        oc, orr = oa.fcos_location_targets(torch.cat(locs).numpy(), soi.numpy(), c[2].numpy(), c[3].numpy(),
                                           [len(l) for l in locs], strides, radius, 80)
        assert np.array_equal(cls[i].cpu().numpy(), oc) and np.array_equal(reg[i].cpu().numpy(), orr)


def test_fcos_location_targets_first_gt_at_origin_shortcut():
    """get_sample_region's `center_x[..., 0].sum() == 0` shortcut: a first GT centred at x == 0 disables centre
    sampling for the whole image (everything background), reproduced as is."""
    from test_oracle_assign import _fcos_case
    from slenderobjdet_b200.targets import compute_targets_for_locations
    locs, soi, boxes, classes, strides = _fcos_case(6)
    boxes[0] = torch.tensor([-20.0, 10.0, 20.0, 60.0])
    cls, reg = compute_targets_for_locations([l.cuda() for l in locs], [(boxes.cuda(), classes.cuda())], soi.cuda(),
                                             strides, 1.5, 80)
    oc, orr = oa.fcos_location_targets(torch.cat(locs).numpy(), soi.numpy(), boxes.numpy(), classes.numpy(),
                                       [len(l) for l in locs], strides, 1.5, 80)
    assert (oc == 80).all() and np.array_equal(cls[0].cpu().numpy(), oc) and np.array_equal(reg[0].cpu().numpy(), orr)


@pytest.mark.parametrize("radius,norm", [(0.0, False), (1.5, True)])
def test_fcos_topk_location_targets_bit_exact(radius, norm):
    """compute_topk_targets_for_locations (fcos/utils.py:215-292): classes, ltrb targets (optionally stride-
    normalised) and the per-GT top-5-by-centerness mask identical to the oracle, on the P3-P7 grid, two images."""
    from test_oracle_assign import _fcos_case
    from slenderobjdet_b200.targets import compute_topk_targets_for_locations
    lv = ((100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128))
    cases = [_fcos_case(s, 60, lv) for s in (8, 9)]
    locs, soi, _, _, strides = cases[0]
    npts = [len(l) for l in locs]
    cls, reg, tk = compute_topk_targets_for_locations([l.cuda() for l in locs], [(c[2].cuda(), c[3].cuda()) for c in cases],
                                                      soi.cuda(), strides, radius, 80, norm_reg_targets=norm, topk=5)
    assert tk.dtype == torch.bool and tk.shape == cls.shape
    for i, c in enumerate(cases):
        oc, orr, idx = oa.fcos_location_targets(torch.cat(locs).numpy(), soi.numpy(), c[2].numpy(), c[3].numpy(), npts,
                                                strides, radius, 80, return_index=True)
        ot = oa.fcos_topk_locations(oc, orr, idx, 80, topk=5)
        if norm:
            orr = orr / np.concatenate([np.full(n, s, np.float32) for n, s in zip(npts, strides)])[:, None]
        assert np.array_equal(cls[i].cpu().numpy(), oc) and np.array_equal(reg[i].cpu().numpy(), orr)
        assert np.array_equal(tk[i].cpu().numpy(), ot) and ot.sum() > 20


def test_fcos_rpd_get_ground_truth_bit_exact():
    """FCOSRepPoints.get_ground_truth (fcos_rpd_s1_topk.py:320-376): both stages against the oracle."""
    from test_oracle_assign import _fcos_case
    from slenderobjdet_b200.targets import fcos_rpd_get_ground_truth
    lv = ((100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128))
    g = torch.Generator().manual_seed(21)
    cases = [_fcos_case(s, 50, lv) for s in (12, 13)]
    locs, soi, _, _, strides = cases[0]
    centers = torch.cat(locs)
    sizes = [(800, 1333), (704, 1216)]
    init = []
    for _ in cases:   # stage-1 boxes: noisy boxes around the centres
        wh = torch.exp(torch.rand(centers.shape[0], 2, generator=g) * 3.5 + 2.0)
        init.append(torch.cat([centers - wh / 2, centers + wh / 2], 1))
    gts = [(c[2].cuda(), c[3].cuda(), sz) for c, sz in zip(cases, sizes)]
    ic, ir, rc, rr, tk = fcos_rpd_get_ground_truth([l.cuda() for l in locs], [b.cuda() for b in init], gts, strides, 0.0, 80)
    npts = [len(l) for l in locs]
    for i, (c, sz) in enumerate(zip(cases, sizes)):
        oc, orr, idx = oa.fcos_location_targets(centers.numpy(), soi.numpy(), c[2].numpy(), c[3].numpy(), npts, strides,
                                                0.0, 80, return_index=True)
        assert np.array_equal(ic[i].cpu().numpy(), oc) and np.array_equal(ir[i].cpu().numpy(), orr)
        assert np.array_equal(tk[i].cpu().numpy(), oa.fcos_topk_locations(oc, orr, idx, 80, topk=5))
        c2, r2 = oa.fcos_rpd_refine_targets(centers.numpy(), init[i].numpy(), c[2].numpy(), c[3].numpy(), sz, 80)
        assert np.array_equal(rc[i].cpu().numpy(), c2) and np.array_equal(rr[i].cpu().numpy(), r2)
        assert (c2 == -1).any() and (c2 == 80).any() and ((c2 >= 0) & (c2 < 80)).any()
