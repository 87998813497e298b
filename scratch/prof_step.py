"""scratch: one P3 DeformConv fwd + bwd through the C ABI (for ncu captures)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scratch.bench_all import run
run(2, 256, 100, 168, 256)
