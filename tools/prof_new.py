"""Profiling driver for the round-2 additions: the kind::tf32 forward (P3 map, both modes), one tower layer (plain
convolution + GroupNorm+ReLU over P3-P7 x 2 towers, batch 2, bf16) forward and backward.

    ncu --set full --clock-control none --import-source on -k regex:"dcn_fwd_tf32|dcn_fwd_tc_kernel|gn_|wgrad_col" \
        -o gpurun_out/r2_prof_new python tools/prof_new.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import slenderobjdet_b200 as sdb  # noqa: E402
import slenderobjdet_b200.layers as L  # noqa: E402

LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 256, 100, 168, generator=g).cuda()
w = (torch.randn(256, 256, 3, 3, generator=g) * 0.01).cuda()
off = (torch.randn(2, 18, 100, 168, generator=g) * 2).cuda()
for mode in ("tf32x3", "tf32"):
    with sdb.dcn_math(mode), torch.no_grad():
        sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
xs = [torch.randn(2, 256, H, W, generator=g).to(bf).cuda().requires_grad_() for _ in range(2) for (H, W) in LEVELS]
gys = [torch.randn(2, 256, H, W, generator=g).to(bf).cuda() for _ in range(2) for (H, W) in LEVELS]
ids = [t for t in range(2) for _ in LEVELS]
ws = [(torch.randn(256, 256, 3, 3, generator=g) * 0.01).to(bf).cuda().requires_grad_() for _ in range(2)]
gam = [torch.ones(256, device="cuda").requires_grad_() for _ in range(2)]
bet = [torch.zeros(256, device="cuda").requires_grad_() for _ in range(2)]
ys = L.conv2d_multi(xs, ws, None, 1, 1, ids)
zs = L.group_norm_relu_multi(ys, gam, bet, 32, 1e-5, ids)
torch.autograd.backward(zs, gys)
torch.cuda.synchronize()
print("done")
