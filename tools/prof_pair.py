"""ncu driver: forward of the benchmarked head with the CTA-pair switch on (argv[1] = 1) or off (0)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from slenderobjdet_b200 import _lib as L  # noqa: E402

L.lib().sdb_set_forward_pair(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
dev = torch.device("cuda", 0)
wl = bench.Workload(torch, L, dev, seed=0, batch=2)
st = torch.cuda.current_stream(dev)
for _ in range(3):
    wl.phase_forward(st)
torch.cuda.synchronize()
