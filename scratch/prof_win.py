import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scratch.bench_fwd import run
print(run(2, 256, 100, 168, 256, iters=1))
