"""Developer check of the CTA-pair (cta_group::2) forward: parity tests with the switch on, then A/B timing.

    python tools/pair_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytest  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from slenderobjdet_b200 import _lib as L  # noqa: E402

lib = L.lib()
lib.sdb_set_forward_pair(int(os.environ.get("PAIR", "0")))
rc = pytest.main(["-x", "-q", "tests/test_gpu_dcn_large.py", "tests/test_gpu_dcn_multi.py", "tests/test_gpu_conv_tower.py",
                  "-k", "forward or benchmarked or whole or plain or towers"])
print("pytest rc", rc)
if rc != 0:
    sys.exit(1)
dev = torch.device("cuda", 0)
wl = bench.Workload(torch, L, dev, seed=0, batch=2)
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for pair in (0, 1, 0, 1):
    lib.sdb_set_forward_pair(pair)
    for _ in range(3):
        wl.phase_forward(st)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wl.phase_forward(st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("pair", pair, "forward phase (prep + pack + kernel) median %.1f us" % ts[len(ts) // 2])
