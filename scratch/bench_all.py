"""scratch: time fwd / bwd_data / bwd_weight (tensor-core path) per FPN level through the C ABI."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slenderobjdet_b200 import _lib

def bench(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def run(N, C, H, W, O, mask=False, sigma=2.0):
    lib = _lib.lib()
    g = _lib.Geom(N, C, H, W, O, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    gp = ctypes.byref(g)
    x = torch.randn(N, C, H, W, device="cuda"); w = torch.randn(O, C, 3, 3, device="cuda") * 0.01
    off = torch.randn(N, 18, H, W, device="cuda") * sigma
    m = torch.rand(N, 9, H, W, device="cuda") if mask else None
    out = torch.empty(N, O, H, W, device="cuda"); gy = torch.randn_like(out)
    gx = torch.zeros_like(x); go = torch.empty_like(off); gm = torch.empty_like(m) if mask else None
    gw = torch.zeros_like(w)
    ws = [torch.empty(max(1, lib.sdb_dcn_workspace_bytes(op, gp, 0, 1)), dtype=torch.uint8, device="cuda") for op in range(3)]
    pk = torch.empty(lib.sdb_dcn_packed_input_bytes(gp, 1), dtype=torch.uint8, device="cuda")
    st = _lib.stream_ptr(); P = _lib.ptr
    f = lambda: _lib.check(lib.sdb_dcn_forward(P(x), P(off), P(m), P(w), None, P(out), gp, 0, 1, P(ws[0]), ws[0].numel(), P(pk), st))
    bd = lambda: _lib.check(lib.sdb_dcn_backward_data(P(x), P(off), P(m), P(w), P(gy), P(gx), P(go), P(gm), gp, 0, 1, P(ws[1]), ws[1].numel(), P(pk), st))
    bw = lambda: _lib.check(lib.sdb_dcn_backward_weight(P(x), P(off), P(m), P(gy), P(gw), None, 1.0, gp, 0, 1, P(ws[2]), ws[2].numel(), P(pk), st))
    tf, tbd, tbw = bench(f), bench(bd), bench(bw)
    fl = 2.0 * N * H * W * O * C * 9 / 1e6
    print(f"N={N} {H}x{W} C={C} O={O} mask={mask}: fwd {tf:8.1f} us ({fl/tf:6.1f} TF/s)  bwd_data {tbd:8.1f} us ({fl/tbd:6.1f})  bwd_weight {tbw:8.1f} us ({fl/tbw:6.1f})", flush=True)
    return tf, tbd, tbw

if __name__ == "__main__":
    tot = [0, 0, 0]
    for (H, W) in [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]:
        r = run(2, 256, H, W, 256)
        tot = [a + b for a, b in zip(tot, r)]
    print("P3-P7 N=2 one DCN: fwd %.1f bwd_data %.1f bwd_weight %.1f us; x2 DCNs fwd+bwd = %.1f us" % (*tot, 2 * sum(tot)))
    run(16, 256, 100, 168, 256)
    run(8, 256, 100, 168, 256, mask=True)
