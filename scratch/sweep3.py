import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scratch.bench_fwd import run
os.environ["SDB_TC_WIN"] = "0"
shapes = [(2, 256, 100, 168, 256), (16, 256, 100, 168, 256)]
for lpp, kb, nsa in [(8, 226, 4), (8, 200, 4), (8, 200, 3), (16, 226, 3), (16, 226, 4), (8, 180, 4)]:
    os.environ["SDB_TC_LPP"], os.environ["SDB_TC_SMEM_KB"], os.environ["SDB_TC_NSA"] = str(lpp), str(kb), str(nsa)
    for s_ in shapes:
        try:
            ms, tf = run(*s_)
            print(f"lpp={lpp} smemKB={kb} nsa={nsa} shape={s_}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
        except Exception as e:
            print(f"lpp={lpp} smemKB={kb} nsa={nsa}: {e}")
