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
