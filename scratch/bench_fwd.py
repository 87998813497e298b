"""scratch: time the tcgen05 forward kernel at the RepPoints head shapes under several tunings."""
import ctypes, os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slenderobjdet_b200 import _lib

def run(N, C, H, W, O, iters=20, mask=False):
    lib = _lib.lib()
    g = _lib.Geom(N, C, H, W, O, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    x = torch.randn(N, C, H, W, device="cuda")
    w = torch.randn(O, C, 3, 3, device="cuda") * 0.01
    base = torch.tensor([(y, x_) for y in (-1, 0, 1) for x_ in (-1, 0, 1)], dtype=torch.float32, device="cuda").view(1, 18, 1, 1)
    off = torch.randn(N, 18, H, W, device="cuda") * float(os.environ.get("SDB_SIGMA", "2.0"))  # sigma=2 px around the regular grid
    m = torch.rand(N, 9, H, W, device="cuda") if mask else None
    out = torch.empty(N, O, H, W, device="cuda")
    wsb = lib.sdb_dcn_workspace_bytes(0, ctypes.byref(g), 0, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    pk = torch.empty(lib.sdb_dcn_packed_input_bytes(ctypes.byref(g), 1), dtype=torch.uint8, device="cuda")
    st = _lib.stream_ptr()
    def call():
        _lib.check(lib.sdb_dcn_forward(_lib.ptr(x), _lib.ptr(off), _lib.ptr(m), _lib.ptr(w), None, _lib.ptr(out),
                                       ctypes.byref(g), 0, 1, _lib.ptr(ws), wsb, _lib.ptr(pk), st))
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * N * H * W * O * C * 9
    return ms, flops / ms / 1e9

if __name__ == "__main__":
    shapes = [(2, 256, 100, 168, 256), (16, 256, 100, 168, 256)]
    for lpp, kb, nsa in [(16, 200, 2), (16, 200, 3), (16, 226, 4), (8, 200, 2), (8, 200, 4), (8, 226, 4), (32, 226, 2), (16, 130, 2), (16, 100, 2)]:
        os.environ["SDB_TC_LPP"], os.environ["SDB_TC_SMEM_KB"], os.environ["SDB_TC_NSA"] = str(lpp), str(kb), str(nsa)
        for s_ in shapes:
            try:
                ms, tf = run(*s_)
                print(f"lpp={lpp} smemKB={kb} nsa={nsa} shape={s_}: {ms*1e3:8.1f} us (incl. pack+wprep)  {tf:7.1f} TFLOP/s", flush=True)
            except Exception as e:
                print(f"lpp={lpp} smemKB={kb} nsa={nsa}: {e}")
