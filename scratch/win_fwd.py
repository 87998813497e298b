"""scratch: forward kernel, window-staged vs L2-gather, with timing-experiment switches"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scratch.bench_fwd import run
shapes = [(2, 256, 100, 168, 256), (16, 256, 100, 168, 256), (2, 256, 50, 84, 256), (2, 256, 7, 11, 256)]
cfgs = [dict(SDB_TC_WIN="0"), dict(SDB_TC_WIN_CL="1"), dict(SDB_TC_WIN_CL="2"), dict(SDB_TC_WIN_CL="2", SDB_TC_WIN_NSB="4"),
        dict(SDB_TC_WIN_CL="2", SDB_TC_DEBUG="1"), dict(SDB_TC_WIN_CL="2", SDB_TC_DEBUG="2"),
        dict(SDB_TC_WIN_CL="1", SDB_TC_DEBUG="2"), dict(SDB_TC_WIN_CL="2", SDB_TC_DEBUG="3")]
if len(sys.argv) > 1:
    cfgs = [eval(a) for a in sys.argv[1:]]
for c in cfgs:
    for k in ("SDB_TC_WIN", "SDB_TC_WIN_NSB", "SDB_TC_DEBUG", "SDB_TC_WIN_R", "SDB_TC_WIN_CL"):
        os.environ.pop(k, None)
    os.environ.update(c)
    for s_ in shapes[:2] if "SDB_TC_DEBUG" in c else shapes:
        try:
            ms, tf = run(*s_)
            print(f"{c} shape={s_}: {ms*1e3:8.1f} us (incl. pack+wprep)  {tf:7.1f} TFLOP/s", flush=True)
        except Exception as e:
            print(f"{c}: {e}")
