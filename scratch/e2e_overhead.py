"""scratch: host enqueue time vs device time of the public-API step (DeformConv fwd + autograd bwd)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slenderobjdet_b200 as sdb
LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
dev = torch.device("cuda", 0); bf = torch.bfloat16
convs = [sdb.DeformConv(256, 256, 3, 1, 1).to(dev, bf) for _ in range(2)]
data = []
for (H, W) in LEVELS:
    data.append(dict(x=[torch.randn(2, 256, H, W, device=dev, dtype=bf) for _ in range(2)],
                     gy=[torch.randn(2, 256, H, W, device=dev, dtype=bf) for _ in range(2)],
                     off=torch.randn(2, 18, H, W, device=dev) * 2))
def step():
    for lv in data:
        off = lv["off"].detach().requires_grad_()
        for b in range(2):
            x = lv["x"][b].detach().requires_grad_()
            y = convs[b](x, off)
            y.backward(lv["gy"][b])
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("enqueue %.2f ms, total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
