import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ["SDB_TC_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "64"
os.environ.setdefault("SDB_TC_WIN_CL", "1")
from scratch.bench_fwd import run
from slenderobjdet_b200 import _lib
print(run(2, 256, 100, 168, 256, iters=1))
torch.cuda.synchronize()
buf = np.zeros(3 * 2 * 64 * 4, dtype=np.uint64)
L = ctypes.CDLL(_lib.lib()._name) if hasattr(_lib.lib(), "_name") else _lib.lib()
print("rc", L.sdb_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size))
t = buf.reshape(3, 2, 64, 4).astype(np.int64)
t0 = t[0, 0, 0, 0]
for k in range(2):
    print("tile", k)
    for st in range(int(sys.argv[2]) if len(sys.argv) > 2 else 36):
        g, m, w = t[0, k, st] - t0, t[1, k, st] - t0, t[2, k, st] - t0
        print(f"st {st:2d} gather: wait {g[0]:7d} got {g[1]:7d} done {g[2]:7d} arrived {g[3]:7d} | mma: wait {m[0]:7d} got {m[1]:7d} issued {m[2]:7d} | w issue {w[0]:7d}")
