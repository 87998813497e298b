"""scratch: the bandwidth-bound secondary kernels (BASELINE.json configs[3]) against their HBM roofline.
Per image: fused IoU + per-anchor max + per-GT top-k assignment (X = 22 400 anchors, M = 100 GT, k = 9);
per batch of 2: fused sigmoid focal loss + gradient over [44 800, 80] logits; fused GIoU loss + gradient over
2 000 positive boxes.  CUDA events, L2 flushed (256 MiB memset) between iterations, public Python API."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slenderobjdet_b200 as sdb
from slenderobjdet_b200 import layers as L

dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return ts[len(ts) // 2]

g = torch.Generator().manual_seed(0)
X, M = 22400, 100
ctr = torch.rand(X, 2, generator=g) * torch.tensor([1333.0, 800.0])
wh = torch.exp(torch.rand(X, 2, generator=g) * 4.2 + 2.0)
anchors = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).to(dev)
c2 = torch.rand(M, 2, generator=g) * torch.tensor([1333.0, 800.0])
wh2 = torch.exp(torch.rand(M, 2, generator=g) * 4.2 + 2.0)
gt = torch.cat([c2 - wh2 / 2, c2 + wh2 / 2], 1).to(dev)
tk = sdb.TopKMatcher([0.3, 0.7], [0, -1, 1], topk=9)
mt = sdb.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=True)
rows = []
us = timeit(lambda: tk.from_boxes(gt, anchors))
by = X * 16 + M * 16 + X * 9
rows.append(("TopKMatcher.from_boxes (fused IoU + max + top-9), 100 x 22 400", us, by))
us = timeit(lambda: mt.from_boxes(gt, anchors))
rows.append(("Matcher.from_boxes (allow_low_quality), 100 x 22 400", us, by))
us = timeit(lambda: sdb.pairwise_iou(gt, anchors))
rows.append(("pairwise_iou 100 x 22 400 (materialised matrix)", us, X * 16 + M * 16 + M * X * 4))

R, K = 2 * X, 80
logits = (torch.randn(R, K, generator=g) * 2 - 4.6).to(dev).requires_grad_()
cls = torch.full((R,), K, dtype=torch.int64)
pos = torch.randperm(R, generator=g)[: R // 100]
cls[pos] = torch.randint(0, K, (pos.numel(),), generator=g)
cls = cls.to(dev)
def focal():
    logits.grad = None
    L.sigmoid_focal_loss_from_class_idx(logits, cls, 0.25, 2.0).backward()
us = timeit(focal)
rows.append(("sigmoid focal loss fwd+grad, [44 800, 80] logits + class index", us, R * K * 4 * 2 + R * 8))

P = 2000
b1 = torch.rand(P, 4, generator=g); b1[:, 2:] += b1[:, :2] + 0.1
b2 = torch.rand(P, 4, generator=g); b2[:, 2:] += b2[:, :2] + 0.1
b1 = (b1 * 100).to(dev).requires_grad_(); b2 = (b2 * 100).to(dev)
def giou():
    b1.grad = None
    L.giou_loss(b1, b2, reduction="sum").backward()
us = timeit(giou)
rows.append(("GIoU loss fwd+grad, 2 000 boxes", us, P * 16 * 3))
print("peak HBM %.0f GB/s (MEASURED_PEAKS.json)" % peaks["hbm_gbs"])
for name, us, by in rows:
    print("%-72s %8.1f us  %8.2f MB algorithmic  %7.1f GB/s  %5.1f %% of HBM peak" % (name, us, by / 1e6, by / us / 1e3, 100 * by / us / 1e3 / peaks["hbm_gbs"]))

# ---- rows a7 / a11 / a12 / a13 / a18: target assignment + glue kernels ------------------------
from slenderobjdet_b200 import targets as T
rows = []
lv = ((100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128))
pts, strides, locs, soi = [], [], [], []
ranges = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, 100000000]]
for (h, w, s), rg in zip(lv, ranges):
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    p = torch.stack([xs.reshape(-1) * s, ys.reshape(-1) * s], 1)
    pts.append(p); strides.append(torch.full((h * w,), float(s))); locs.append((p + s // 2).to(dev))
    soi.append(torch.tensor(rg, dtype=torch.float32)[None].expand(h * w, -1))
pts, strides, soi = torch.cat(pts).to(dev), torch.cat(strides).to(dev), torch.cat(soi).to(dev)
glab = torch.randint(0, 80, (M,), generator=g).to(dev)
us = timeit(lambda: T.point_targets(pts, strides, gt, glab, 80))
rows.append(("point_targets (a11), 22 400 points x 100 GT", us, X * 12 + M * 24 + X * 24))
cand = anchors.clone()
us = timeit(lambda: T.bbox_targets(cand, gt, glab, 80))
rows.append(("bbox_targets (a12), 22 400 boxes x 100 GT", us, X * 16 + M * 24 + X * 24))
for radius in (0.0, 1.5):
    us = timeit(lambda: T.compute_targets_for_locations(locs, [(gt, glab)], soi, [8, 16, 32, 64, 128], radius, 80))
    rows.append(("fcos compute_targets_for_locations (a13), radius %.1f, 1 image" % radius, us, X * 16 + M * 24 + X * 24))
ltrb = torch.rand(2 * X, 4, generator=g).to(dev) * 100 + 1
us = timeit(lambda: L.compute_centerness_targets(ltrb))
rows.append(("compute_centerness_targets (a18), 44 800 rows", us, 2 * X * 20))
p3 = torch.randn(2, 18, 100, 168, generator=g).to(dev).requires_grad_()
gop = torch.randn(2, 18, 100, 168, generator=g).to(dev)
def offs():
    p3.grad = None
    L.reppoints_dcn_offset(p3, 0.1).backward(gop)
us = timeit(offs)
rows.append(("reppoints_dcn_offset fwd+bwd (a7), P3 N=2", us, 2 * 18 * 16800 * 4 * 4))
for name, us, by in rows:
    print("%-72s %8.1f us  %8.2f MB algorithmic  %7.1f GB/s" % (name, us, by / 1e6, by / us / 1e3))
