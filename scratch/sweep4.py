import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scratch.bench_fwd import run
os.environ["SDB_TC_WIN"] = "0"
shapes = [(2, 256, 100, 168, 256), (16, 256, 100, 168, 256)]
for lpp, kb, nsa, carve in [(16, 200, 2, None), (8, 135, 2, 60), (8, 135, 2, 100), (8, 120, 2, 55), (16, 165, 2, 75), (8, 165, 3, 75)]:
    os.environ["SDB_TC_LPP"], os.environ["SDB_TC_SMEM_KB"], os.environ["SDB_TC_NSA"] = str(lpp), str(kb), str(nsa)
    if carve is None: os.environ.pop("SDB_TC_CARVEOUT", None)
    else: os.environ["SDB_TC_CARVEOUT"] = str(carve)
    for s_ in shapes:
        try:
            ms, tf = run(*s_)
            print(f"lpp={lpp} smemKB={kb} nsa={nsa} carveout={carve} shape={s_}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
        except Exception as e:
            print(f"lpp={lpp} smemKB={kb} nsa={nsa} carve={carve}: {e}")
