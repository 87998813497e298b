"""Forward of ONE DeformConv 256->256 3x3 through the public operator API in every math mode (float32 tensors; bf16
mode also with bf16 tensors), CUDA-event times after warm-up, and the relative error against the fp32 SIMT result.

    python tools/time_fwd_modes.py [--batch 2] [--hw 100 168]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import slenderobjdet_b200 as sdb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--hw", type=int, nargs=2, default=[100, 168])
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
H, W = a.hw
g = torch.Generator().manual_seed(0)
x = torch.randn(a.batch, 256, H, W, generator=g).cuda()
w = (torch.randn(256, 256, 3, 3, generator=g) * 0.01).cuda()
off = (torch.randn(a.batch, 18, H, W, generator=g) * 2).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for mode, dt in (("fp32", torch.float32), ("tf32x3", torch.float32), ("tf32", torch.float32), ("bf16", torch.float32),
                 ("bf16", torch.bfloat16)):
    xx, ww = x.to(dt), w.to(dt)
    with sdb.dcn_math(mode), torch.no_grad():
        for _ in range(3):
            y = sdb.deform_conv(xx, off, ww, 1, 1, 1, 1, 1)
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = sdb.deform_conv(xx, off, ww, 1, 1, 1, 1, 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    y = y.float()
    if ref is None:
        ref = y.double()
    err = float((y.double() - ref).norm() / ref.norm())
    flops = 2.0 * a.batch * H * W * 256 * 2304
    print("%-7s %-9s median %8.1f us  min %8.1f us  %7.1f TFLOP/s  rel err vs fp32 %.2e"
          % (mode, str(dt).split(".")[1], ts[len(ts) // 2], ts[0], flops / ts[len(ts) // 2] / 1e6, err))
