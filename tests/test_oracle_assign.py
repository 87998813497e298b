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
