"""scratch: run-to-run spread of grad_input's error on the adversarial all-taps-collide case"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_dcn as T
from conftest import rel_err
for spread in (0.0, 0.6):
    N, C, H, W, O = 2, 64, 11, 13, 64
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, C, H, W, generator=g); w = torch.randn(O, C, 3, 3, generator=g) * 0.05
    gy = torch.randn(N, O, H, W, generator=g); off = torch.zeros(N, 18, H, W)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    for i in range(3):
        for j in range(3):
            t = i * 3 + j
            off[:, 2 * t] = (H / 2 + 0.3) - (ys - 1 + i); off[:, 2 * t + 1] = (W / 2 + 0.6) - (xs - 1 + j)
    off += torch.randn(N, 18, H, W, generator=g) * spread
    c = dict(x=x.numpy(), offset=off.numpy(), weight=w.numpy(), grad_out=gy.numpy(), cfg=np.array([1, 1, 1, 1, 1, 1, 1, 1]))
    yo, go = T._oracle(c)
    errs = []
    for _ in range(40):
        y, xd, offd, wd, _, _ = T._run(c, "bf16")
        errs.append(rel_err(xd.grad.cpu().numpy(), go["grad_x"]))
    errs = np.array(errs)
    print("spread", spread, "grad_x rel err: min %.4g median %.4g max %.4g" % (errs.min(), np.median(errs), errs.max()))
