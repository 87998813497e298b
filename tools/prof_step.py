"""Profiling driver: N steps of the benchmarked workload (bench.Workload) through the C ABI, eager launches.

    ncu --set full --clock-control none --import-source on -k regex:dcn_ -s 8 -c 4 -o gpurun_out/prof python tools/prof_step.py
    ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of 2 steps> -c <launches of 1 step> --csv \
        --log-file gpurun_out/launches.csv python tools/prof_step.py --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from slenderobjdet_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--levels", default="all", help="'all' or e.g. 'P3'")
ap.add_argument("--modulated", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
levels = bench.LEVELS if a.levels == "all" else [bench.LEVELS[int(a.levels[1]) - 3]]
wl = bench.Workload(torch, L, dev, seed=0, batch=a.batch, levels=levels, modulated=a.modulated)
st = torch.cuda.current_stream(dev)
for _ in range(a.steps):
    wl.step(st)
torch.cuda.synchronize()
print("launches per step:", L.lib().sdb_launch_count() // a.steps)
