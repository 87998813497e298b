import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scratch.bench_fwd import run
shape = (16, 256, 100, 168, 256)
for dbg in (3, 7, 4, 0):
    os.environ["SDB_TC_DEBUG"] = str(dbg)
    ms, tf = run(*shape)
    print(f"dbg={dbg} (1=no weights/MMA, 2=no gather, 4=no epilogue stores) {shape}: {ms*1e3:8.1f} us", flush=True)
os.environ["SDB_TC_DEBUG"] = "0"
for t in ("8x16", "4x32", "2x64", "16x8"):
    os.environ["SDB_TC_TILE"] = t
    ms, tf = run(*shape)
    print(f"tile={t} {shape}: {ms*1e3:8.1f} us", flush=True)
