"""Measurement: forward / backward kernel times of the benchmarked workload under the hidden tuning overrides
(sdb_debug_tune: lanes per pixel of the forward gather, shared-memory budget = L1 size, A stages)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from slenderobjdet_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
lib = L.lib()
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cfgs = [(0, 0, 0), (16, 170, 2), (8, 200, 2), (8, 170, 2), (8, 140, 2), (8, 172, 3), (8, 156, 3)]
for batch in (2, 16):
    for lpp, kb, nsa in cfgs:
        lib.sdb_debug_tune(lpp, kb, nsa)
        try:
            wl = bench.Workload(torch, L, dev, seed=0, batch=batch)
            for _ in range(2):
                wl.step(st)
            torch.cuda.synchronize()
            lib.sdb_profile_reset(); lib.sdb_profile_enable(1)
            for _ in range(5):
                flush.zero_()
                wl.step(st)
            lib.sdb_profile_enable(0)
            torch.cuda.synchronize()
            out = []
            for slot in range(4):
                ms, n = ctypes.c_float(0), ctypes.c_int(0)
                lib.sdb_profile_read(slot, ctypes.byref(ms), ctypes.byref(n))
                out.append(ms.value / 5 * 1e3)
            print("batch %d lpp %2d smem %3d nsa %d : fwd %7.1f goff %7.1f wgrad %7.1f dx %7.1f us" % ((batch, lpp, kb, nsa) + tuple(out)), flush=True)
            del wl
        except Exception as e:
            print("batch %d lpp %d smem %d nsa %d : FAILED %r" % (batch, lpp, kb, nsa, e), flush=True)
lib.sdb_debug_tune(0, 0, 0)
