"""Tower path timings on the benchmarked head shapes: P3-P7 of an 800x1344 image, batch 2, two
towers, 256 channels, bf16.  One tower LAYER = conv 3x3 -> GroupNorm(32) -> ReLU over 5 levels x 2 towers.
Prints CUDA-event medians of CUDA-graph replays (no host launch overhead; L2 flushed between iterations) for this library and, beside it, for the eager PyTorch
layers the reference runs (cuDNN convolution + native group_norm + relu, same tensors) as the library baseline.

    python tools/time_towers.py [--batch 2] [--iters 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import slenderobjdet_b200.layers as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
if os.environ.get("PAIR"):   # developer switch: CTA-pair forward kernel
    from slenderobjdet_b200 import _lib as _L
    _L.lib().sdb_set_forward_pair(int(os.environ["PAIR"]))
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
C = 256
xs = [torch.randn(a.batch, C, H, W, generator=g).to(bf).cuda().requires_grad_() for _ in range(2) for (H, W) in LEVELS]
gys = [torch.randn(a.batch, C, H, W, generator=g).to(bf).cuda() for _ in range(2) for (H, W) in LEVELS]
ids = [t for t in range(2) for _ in LEVELS]
ws = [(torch.randn(C, C, 3, 3, generator=g) * 0.01).to(bf).cuda().requires_grad_() for _ in range(2)]
gam = [torch.ones(C, device="cuda", dtype=bf).requires_grad_() for _ in range(2)]
bet = [torch.zeros(C, device="cuda", dtype=bf).requires_grad_() for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
px = sum(H * W for (H, W) in LEVELS) * a.batch * 2
flop_pass = 2.0 * px * C * C * 9


def timed(fn, iters=a.iters):
    """median time of one CUDA-graph replay of fn (no host launch overhead on either side of the comparison)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            fn()
    torch.cuda.synchronize()
    fn = graph.replay
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def ours_conv_fwd():
    with torch.no_grad():
        return L.conv2d_multi(xs, ws, None, 1, 1, ids)


def ours_conv_fwd_bwd():
    ys = L.conv2d_multi(xs, ws, None, 1, 1, ids)
    torch.autograd.backward(ys, gys)


xcl = [x.detach().contiguous(memory_format=torch.channels_last).requires_grad_() for x in xs]
gcl = [t.contiguous(memory_format=torch.channels_last) for t in gys]
wcl = [w.detach().contiguous(memory_format=torch.channels_last).requires_grad_() for w in ws]


def torch_conv_fwd():
    with torch.no_grad():
        return [F.conv2d(xcl[i], wcl[ids[i]], None, 1, 1) for i in range(len(xcl))]


def torch_conv_fwd_bwd():
    ys = [F.conv2d(xcl[i], wcl[ids[i]], None, 1, 1) for i in range(len(xcl))]
    torch.autograd.backward(ys, gcl)


def ours_gn_fwd_bwd():
    ys = L.group_norm_relu_multi(xs, gam, bet, 32, 1e-5, ids)
    torch.autograd.backward(ys, gys)


def torch_gn_fwd_bwd():
    ys = [F.relu(F.group_norm(xs[i], 32, gam[ids[i]], bet[ids[i]], 1e-5)) for i in range(len(xs))]
    torch.autograd.backward(ys, gys)


def ours_gn_fwd():
    with torch.no_grad():
        return L.group_norm_relu_multi(xs, gam, bet, 32, 1e-5, ids)


def torch_gn_fwd():
    with torch.no_grad():
        return [F.relu(F.group_norm(xs[i], 32, gam[ids[i]], bet[ids[i]], 1e-5)) for i in range(len(xs))]


act_bytes = px * C * 2
rows = [("conv 3x3 forward", ours_conv_fwd, torch_conv_fwd, flop_pass, None),
        ("conv 3x3 forward + backward", ours_conv_fwd_bwd, torch_conv_fwd_bwd, 3 * flop_pass, None),
        ("GroupNorm+ReLU forward", ours_gn_fwd, torch_gn_fwd, None, 3 * act_bytes),
        ("GroupNorm+ReLU forward + backward", ours_gn_fwd_bwd, torch_gn_fwd_bwd, None, 8 * act_bytes)]
for name, ours, ref, flops, nbytes in rows:
    t1, t2 = timed(ours), timed(ref)
    extra = ("%.0f TFLOP/s" % (flops / t1 / 1e6)) if flops else ("%.0f GB/s of %d MB algorithmic" % (nbytes / t1 / 1e3, nbytes >> 20))
    print("%-36s ours %8.1f us (%s)   eager torch %8.1f us   x%.2f" % (name, t1, extra, t2, t2 / t1))
