"""scratch: host time of each C-ABI call (P7-sized problem so the GPU never back-pressures the launch queue)"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slenderobjdet_b200 import _lib
lib = _lib.lib()
N, C, H, W, O = 2, 256, 7, 11, 256
g = _lib.Geom(N, C, H, W, O, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1); gp = ctypes.byref(g)
bf = torch.bfloat16
x = torch.randn(N, C, H, W, device="cuda", dtype=bf); w = (torch.randn(O, C, 3, 3, device="cuda") * 0.01).to(bf)
off = torch.randn(N, 18, H, W, device="cuda") * 2
out = torch.empty(N, O, H, W, device="cuda", dtype=bf); gy = torch.randn_like(out)
gx = torch.zeros_like(x); go = torch.empty_like(off); gw = torch.zeros(O, C, 3, 3, device="cuda")
ws = [torch.empty(max(1, lib.sdb_dcn_workspace_bytes(op, gp, 1, 1)), dtype=torch.uint8, device="cuda") for op in range(3)]
pk = torch.empty(lib.sdb_dcn_packed_input_bytes(gp, 1), dtype=torch.uint8, device="cuda")
st = _lib.stream_ptr(); P = _lib.ptr
calls = {
 "forward": lambda: lib.sdb_dcn_forward(P(x), P(off), None, P(w), None, P(out), gp, 1, 1, P(ws[0]), ws[0].numel(), P(pk), st),
 "backward_data": lambda: lib.sdb_dcn_backward_data(P(x), P(off), None, P(w), P(gy), P(gx), P(go), None, gp, 1, 1, P(ws[1]), ws[1].numel(), P(pk), st),
 "backward_weight": lambda: lib.sdb_dcn_backward_weight(P(x), P(off), None, P(gy), P(gw), None, 1.0, gp, 1, 1, P(ws[2]), ws[2].numel(), P(pk), st),
}
for name, fn in calls.items():
    for _ in range(20): fn()
    torch.cuda.synchronize()
    n0 = lib.sdb_launch_count()
    t0 = time.perf_counter()
    for _ in range(50): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-16s host %.1f us/call (%d launches), incl. drain %.1f us/call" % (name, (t1 - t0) / 50 * 1e6, (lib.sdb_launch_count() - n0) // 50, (t2 - t0) / 50 * 1e6))
import slenderobjdet_b200 as sdb
conv = sdb.DeformConv(C, O, 3, 1, 1).to("cuda", bf)
def pyfwdbwd():
    xx = x.detach().requires_grad_(); oo = off.detach().requires_grad_()
    y = conv(xx, oo); y.backward(gy)
for _ in range(10): pyfwdbwd()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): pyfwdbwd()
t1 = time.perf_counter(); torch.cuda.synchronize()
print("python DeformConv fwd+bwd host %.1f us/call" % ((t1 - t0) / 50 * 1e6))
