"""Developer check of the CTA-pair (cta_group::2) kernels: A/B timing of the forward and backward phases of the benchmarked
head with the pair switches off / on (parity is covered by the GPU test suite, which runs with the defaults = on).

    python tools/pair_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from slenderobjdet_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = torch.device("cuda", 0)
wl = bench.Workload(torch, L, dev, seed=0, batch=2)
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for fp, bp in ((0, 0), (1, 0), (0, 1), (1, 1), (0, 0), (1, 1)):
    lib.sdb_set_forward_pair(fp)
    lib.sdb_set_backward_pair(bp)
    wl.phase_forward(st)
    print("forward pair %d backward pair %d: forward phase %.1f us, backward phase %.1f us (eager launches)"
          % (fp, bp, timed(lambda: wl.phase_forward(st)), timed(lambda: wl.phase_backward(st))))
