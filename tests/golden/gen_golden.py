"""Generate the committed golden fixtures under tests/golden/ (run in the authoring container only).

    python tests/golden/gen_golden.py

Needs /root/reference (read-only) and torchvision; NEITHER is read at test time -- the tests
only load the .npz files written here.  What is generated, and from what:

  dcn_known_answer.npz   the reference's only DCN known-answer test
                         (/root/reference/tests/test_deformable_conv.py:69-87): its inputs, and
                         the expected outputs computed by EXECUTING the reference's own
                         ``my_conv`` / ``my_dconv`` helpers (extracted from that file with ``ast``
                         at generation time; no reference source is copied into this repo).
  dcn_cases.npz          seeded random DCN v1/v2 cases (fractional offsets, borders, groups,
                         deformable groups, stride, dilation, mask+bias) with outputs and all
                         gradients from torchvision.ops.deform_conv2d on CPU (the "CPU
                         deform_conv2d path" BASELINE.json names; the reference's own DeformConv
                         raises NotImplementedError on CPU, deform_conv.py:48-49).
  assign_cases.npz       seeded IoU / Matcher / TopKMatcher cases with outputs from the
                         reference's own Python (boxes.py::pairwise_iou, matcher.py::Matcher,
                         topk_matcher.py::TopKMatcher, loaded by file path), plus the upstream
                         known answers (test_matcher.py:19-27, test_boxes.py:151-173).
  loss_cases.npz         seeded loss cases with outputs from the reference's iou_loss.py /
                         smooth_l1_loss_with_weight.py (by file path) and, for the fvcore
                         formulas, torchvision.ops.sigmoid_focal_loss / generalized_box_iou_loss.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torchvision

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load_by_path(name, path, stubs=None):
    for k, v in (stubs or {}).items():
        sys.modules[k] = v
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _nonzero_tuple(x):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


def load_reference():
    d2 = _stub_module("detectron2")
    d2_layers = _stub_module("detectron2.layers", nonzero_tuple=_nonzero_tuple)
    concern = _stub_module("concern")
    support = _stub_module("concern.support", make_dual=lambda v: v if isinstance(v, tuple) else (v, v))
    stubs = {"detectron2": d2, "detectron2.layers": d2_layers, "concern": concern,
             "concern.support": support}
    ref = types.SimpleNamespace()
    ref.topk = _load_by_path("ref_topk", f"{REF}/slender_det/modeling/matchers/topk_matcher.py", stubs)
    ref.matcher = _load_by_path("ref_matcher", f"{REF}/detectron2/detectron2/modeling/matcher.py", stubs)
    ref.boxes = _load_by_path("ref_boxes", f"{REF}/detectron2/detectron2/structures/boxes.py", stubs)
    ref.iou_loss = _load_by_path("ref_iou_loss", f"{REF}/slender_det/layers/iou_loss.py", stubs)
    ref.sl1 = _load_by_path("ref_sl1", f"{REF}/slender_det/layers/smooth_l1_loss_with_weight.py", stubs)
    ref.grid = _load_by_path("ref_grid", f"{REF}/slender_det/modeling/grid_generator.py", stubs)
    return ref


def gen_dcn_known_answer(ref):
    """Run the reference test's own helper functions on the reference test's own inputs."""
    src = open(f"{REF}/tests/test_deformable_conv.py").read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("my_dconv", "my_conv")]
    ns = {"torch": torch, "F": torch.nn.functional, "uniform_grid": ref.grid.uniform_grid,
          "zero_center_grid": ref.grid.zero_center_grid}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "ref_test_helpers", "exec"), ns)
    # inputs exactly as test_deformable_conv.py:69-81 builds them (CPU tensors)
    weight = torch.arange(9).float().reshape(1, 1, 3, 3).repeat(1, 2, 1, 1)
    grid = ref.grid.uniform_grid(4).unsqueeze(0).permute(0, 3, 1, 2)
    grid = torch.stack([grid[:, 0], torch.zeros_like(grid[:, 0]) + 0.1], 1)
    offsets_1 = ref.grid.zero_center_grid(3).reshape(1, -1, 1, 1).repeat(1, 1, 4, 4)
    offsets_2 = torch.zeros_like(offsets_1)
    exp_conv = ns["my_conv"](grid, weight)
    exp_d2 = ns["my_dconv"](grid, offsets_2, weight)
    exp_d1 = ns["my_dconv"](grid, offsets_1, weight)
    np.savez(os.path.join(OUT, "dcn_known_answer.npz"), x=grid.numpy(), weight=weight.numpy(),
             offsets_1=offsets_1.numpy(), offsets_2=offsets_2.numpy(), expected_conv=exp_conv.numpy(),
             expected_dconv_zero=exp_d2.numpy(), expected_dconv_grid=exp_d1.numpy())


DCN_CASES = [
    # name, N, C, H, W, O, KH, KW, stride, pad, dil, groups, dg, modulated, bias, offset_sigma
    ("v1_basic", 2, 8, 9, 11, 6, 3, 3, 1, 1, 1, 1, 1, False, False, 0.7),
    ("v1_big_offsets", 1, 4, 7, 6, 4, 3, 3, 1, 1, 1, 1, 1, False, False, 5.0),
    ("v1_groups_dg", 2, 8, 8, 8, 4, 3, 3, 1, 1, 1, 2, 2, False, False, 1.0),
    ("v1_stride2_dil2", 1, 4, 11, 13, 3, 3, 3, 2, 2, 2, 1, 1, False, False, 1.5),
    ("v1_k1", 1, 4, 6, 5, 2, 1, 1, 1, 0, 1, 1, 1, False, False, 1.0),
    ("v1_k5x3", 1, 2, 8, 9, 2, 5, 3, (1, 1), (2, 1), (1, 1), 1, 1, False, False, 1.0),
    ("v1_c64", 1, 64, 10, 12, 32, 3, 3, 1, 1, 1, 1, 1, False, False, 2.0),
    ("v2_basic", 2, 8, 9, 11, 6, 3, 3, 1, 1, 1, 1, 1, True, True, 0.7),
    ("v2_nobias_dg2", 1, 8, 7, 9, 4, 3, 3, 1, 1, 1, 1, 2, True, False, 2.0),
    ("v2_c64", 1, 64, 10, 12, 64, 3, 3, 1, 1, 1, 1, 1, True, True, 2.0),
]


def gen_dcn_cases():
    from torchvision.ops import deform_conv2d
    out = {}
    for ci, (name, N, C, H, W, O, KH, KW, st, pd, dl, g, dg, mod, bias, sig) in enumerate(DCN_CASES):
        gen = torch.Generator().manual_seed(1000 + ci)
        pair = lambda v: (v, v) if isinstance(v, int) else v
        st2, pd2, dl2 = pair(st), pair(pd), pair(dl)
        Ho = (H + 2 * pd2[0] - (dl2[0] * (KH - 1) + 1)) // st2[0] + 1
        Wo = (W + 2 * pd2[1] - (dl2[1] * (KW - 1) + 1)) // st2[1] + 1
        x = torch.randn(N, C, H, W, generator=gen)
        w = torch.randn(O, C // g, KH, KW, generator=gen) * 0.2
        off = torch.randn(N, dg * 2 * KH * KW, Ho, Wo, generator=gen) * sig
        m = torch.sigmoid(torch.randn(N, dg * KH * KW, Ho, Wo, generator=gen)) if mod else None
        b = torch.randn(O, generator=gen) if bias else None
        gy = torch.randn(N, O, Ho, Wo, generator=gen)
        leaves = [t for t in (x, off, w, m, b) if t is not None]
        for t in leaves:
            t.requires_grad_(True)
        y = deform_conv2d(x, off, w, b, stride=st2, padding=pd2, dilation=dl2, mask=m)
        y.backward(gy)
        rec = dict(x=x, offset=off, weight=w, grad_out=gy, out=y, grad_x=x.grad, grad_offset=off.grad,
                   grad_weight=w.grad)
        if mod:
            rec.update(mask=m, grad_mask=m.grad)
        if bias:
            rec.update(bias=b, grad_bias=b.grad)
        for k, v in rec.items():
            out[f"{name}/{k}"] = v.detach().numpy()
        out[f"{name}/cfg"] = np.array([st2[0], st2[1], pd2[0], pd2[1], dl2[0], dl2[1], g, dg], np.int64)
    np.savez_compressed(os.path.join(OUT, "dcn_cases.npz"), **out)


def _rand_boxes(gen, n, wmin=8.0, wmax=512.0, img=(800.0, 1333.0)):
    """GT-like boxes: centres uniform in the image, log-uniform size, aspect log-uniform in [1/8, 8]."""
    cy = torch.rand(n, generator=gen) * img[0]
    cx = torch.rand(n, generator=gen) * img[1]
    s = torch.exp(torch.rand(n, generator=gen) * (np.log(wmax) - np.log(wmin)) + np.log(wmin))
    a = torch.exp((torch.rand(n, generator=gen) * 2 - 1) * np.log(8.0) / 2)
    w, h = s * a, s / a
    return torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).float()


def _grid_anchors(size, strides=(8, 16, 32, 64, 128), img=(96, 160)):
    out = []
    for s in strides:
        hs, ws = -(-img[0] // s), -(-img[1] // s)
        ys, xs = torch.meshgrid(torch.arange(hs) * s + s // 2, torch.arange(ws) * s + s // 2, indexing="ij")
        c = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float()
        half = size * s / 2.0
        out.append(torch.cat([c - half, c + half], 1))
    return torch.cat(out, 0)


def gen_assign_cases(ref):
    out = {}
    Boxes = ref.boxes.Boxes
    # upstream known answers (data transcribed from the unit tests)
    out["ka_matcher/q"] = np.array([[0.15, 0.45, 0.2, 0.6], [0.3, 0.65, 0.05, 0.1], [0.05, 0.4, 0.25, 0.4]], np.float32)
    out["ka_matcher/matches"] = np.array([1, 1, 2, 0], np.int64)
    out["ka_matcher/labels"] = np.array([-1, 1, 0, 1], np.int8)
    out["ka_matcher/thresholds"] = np.array([0.3, 0.7], np.float32)  # d2 MODEL.RPN.IOU_THRESHOLDS
    out["ka_matcher/label_values"] = np.array([0, -1, 1], np.int8)   # d2 MODEL.RPN.IOU_LABELS
    b1 = np.array([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]], np.float32)
    b2 = np.array([[0, 0, 1, 1], [0, 0, .5, 1], [0, 0, 1, .5], [0, 0, .5, .5], [.5, .5, 1, 1], [.5, .5, 1.5, 1.5]], np.float32)
    out["ka_iou/boxes1"], out["ka_iou/boxes2"] = b1, b2
    out["ka_iou/expected"] = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]] * 2, np.float32)
    # the reference's own pairwise_iou must reproduce the upstream expectation
    assert torch.allclose(ref.boxes.pairwise_iou(Boxes(torch.tensor(b1)), Boxes(torch.tensor(b2))),
                          torch.tensor(out["ka_iou/expected"]))
    m0, l0 = ref.matcher.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(torch.tensor(out["ka_matcher/q"]))
    assert m0.tolist() == [1, 1, 2, 0] and l0.tolist() == [-1, 1, 0, 1]

    cases = [("small", 7, 4.0, (96, 160)), ("mid", 23, 4.0, (256, 320)), ("slender", 40, 8.0, (800, 1344))]
    for ci, (name, M, asize, img) in enumerate(cases):
        gen = torch.Generator().manual_seed(2000 + ci)
        gt = _rand_boxes(gen, M, 8.0, min(img) / 1.5, img=(float(img[0]), float(img[1])))
        anchors = _grid_anchors(asize, img=img)
        # jitter so no two IoUs tie exactly
        anchors = anchors + (torch.rand(anchors.shape, generator=gen) - 0.5) * 0.37
        iou = ref.boxes.pairwise_iou(Boxes(gt), Boxes(anchors))
        out[f"{name}/gt"], out[f"{name}/anchors"], out[f"{name}/iou"] = gt.numpy(), anchors.numpy(), iou.numpy()
        for k in (1, 9, 10):
            m, l = ref.topk.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=k)(iou)
            s = torch.sort(iou, dim=1, descending=True, stable=True)[0]
            out[f"{name}/topk{k}_matches"], out[f"{name}/topk{k}_labels"] = m.numpy(), l.numpy()
            out[f"{name}/topk{k}_tiefree"] = np.array(bool((s[:, k - 1] != s[:, k]).all()))
        for alq in (False, True):
            m, l = ref.matcher.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=alq)(iou)
            out[f"{name}/matcher{int(alq)}_matches"], out[f"{name}/matcher{int(alq)}_labels"] = m.numpy(), l.numpy()
    # empty-GT behaviour (topk_matcher.py:53-63)
    m, l = ref.topk.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=9)(torch.zeros(0, 11))
    out["empty/matches"], out["empty/labels"] = m.numpy(), l.numpy()
    np.savez_compressed(os.path.join(OUT, "assign_cases.npz"), **out)


def gen_loss_cases(ref):
    from torchvision.ops import sigmoid_focal_loss, generalized_box_iou_loss
    out = {}
    gen = torch.Generator().manual_seed(3000)
    R, K = 257, 80
    logits = (torch.randn(R, K, generator=gen) * 2 - 4.6).requires_grad_(True)
    cls = torch.full((R,), K, dtype=torch.int64)
    pos = torch.randperm(R, generator=gen)[:40]
    cls[pos] = torch.randint(0, K, (40,), generator=gen)
    t = torch.zeros(R, K)
    t[pos, cls[pos]] = 1
    loss = sigmoid_focal_loss(logits, t, alpha=0.25, gamma=2.0, reduction="sum")
    loss.backward()
    out["focal/logits"], out["focal/cls"] = logits.detach().numpy(), cls.numpy()
    out["focal/loss"], out["focal/grad"] = loss.detach().numpy(), logits.grad.numpy()

    P = 193
    for form in ("ltrb", "xyxy"):
        if form == "ltrb":
            pred = (torch.rand(P, 4, generator=gen) * 60 + 0.5)
            tgt = (torch.rand(P, 4, generator=gen) * 60 + 0.5)
            fn = ref.iou_loss.iou_loss
        else:
            pred = _rand_boxes(gen, P, 20, 200)
            # keep every pair overlapping: -log(iou) is NaN in the reference itself otherwise
            side = torch.minimum(pred[:, 2] - pred[:, 0], pred[:, 3] - pred[:, 1])
            tgt = pred + (torch.rand(P, 4, generator=gen) - 0.5) * 0.3 * side[:, None]
            fn = ref.iou_loss.box_iou_loss
        wgt = torch.rand(P, generator=gen)
        for lt in ("iou", "linear_iou", "giou"):
            for use_w in (False, True):
                p = pred.clone().requires_grad_(True)
                l = fn(p, tgt, wgt if use_w else None, loss_type=lt)
                l.backward()
                key = f"{form}_{lt}_{int(use_w)}"
                out[f"{key}/loss"], out[f"{key}/grad"] = l.detach().numpy(), p.grad.numpy()
        out[f"{form}/pred"], out[f"{form}/target"], out[f"{form}/weight"] = pred.numpy(), tgt.numpy(), wgt.numpy()

    pred = torch.randn(P, 4, generator=gen)
    tgt = pred + torch.randn(P, 4, generator=gen) * 0.2
    wgt = torch.rand(P, generator=gen)
    for beta in (0.11, 0.0):
        for use_w in (False, True):
            p = pred.clone().requires_grad_(True)
            l = ref.sl1.smooth_l1_loss_with_weight(p, tgt, wgt if use_w else None, beta, reduction="sum")
            l.backward()
            key = f"sl1_{beta}_{int(use_w)}"
            out[f"{key}/loss"], out[f"{key}/grad"] = l.detach().numpy(), p.grad.numpy()
    out["sl1/pred"], out["sl1/target"], out["sl1/weight"] = pred.numpy(), tgt.numpy(), wgt.numpy()

    b = _rand_boxes(gen, P, 8, 200)
    t2 = b + torch.randn(P, 4, generator=gen) * 6
    p = b.clone().requires_grad_(True)
    l = generalized_box_iou_loss(p, t2, reduction="sum", eps=1e-7)
    l.backward()
    out["giou/pred"], out["giou/target"] = b.numpy(), t2.numpy()
    out["giou/loss"], out["giou/grad"] = l.detach().numpy(), p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "loss_cases.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(4)
    ref = load_reference()
    gen_dcn_known_answer(ref)
    gen_dcn_cases()
    gen_assign_cases(ref)
    gen_loss_cases(ref)
    print("golden fixtures written to", OUT, "torch", torch.__version__, "torchvision", torchvision.__version__)
