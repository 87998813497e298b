import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scratch.bench_all import run
import scratch.bench_all as ba
ba.bench = lambda fn, iters=1: (fn(), torch.cuda.synchronize(), 0.0)[2] + 1.0
for (N, H, W) in [(2, 100, 168), (2, 7, 11), (16, 100, 168)]:
    run(N, 256, H, W, 256)
