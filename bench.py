#!/usr/bin/env python
"""bench.py -- RepPoints-head DCN forward+backward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): RepPoints-v1 R-50-FPN head, 800x1333 input padded to 800x1344,
FPN levels P3-P7 (100x168 ... 7x11, 22 400 points / image), batch 2 per GPU, the head's two
DeformConv(256,256,3,1,1) (cls + pts-refine branch, weights shared across levels, both consuming the level's ONE
dcn_offset, reppointsv2.py:744-748) applied to every level, forward + backward (grad_input, grad_offset, grad_weight),
bf16 tensors, fp32 accumulate.  One "step" = that whole pass (20 deformable convolutions fwd+bwd) INCLUDING the
re-layout of the two weight tensors into the kernels' operand images (weights change every optimiser step).
Synthetic data: randn features, weights randn*0.01, offsets randn*2 px.

`value`  : images/s with every input resident in HBM; the step is three C-ABI calls over a 10-row problem table
           (sdb_dcn_prepare_weights x2, sdb_dcn_forward_multi, sdb_dcn_backward_multi: one launch per kernel over all
           levels and both branches), replayed as a CUDA graph, timed with CUDA events on the launching stream, max
           over ranks.
`e2e`    : images/s through the public Python operator API (slenderobjdet_b200.deform_conv_multi + autograd) with all
           inputs in pinned HOST memory: every step's H2D copies of features / offsets / grad_out and the D2H read
           of its gradients are inside the timed region (uploads of step i+1 overlap the kernels of step i, as a
           training loop's prefetcher does).
N > 1    : one process per GPU, same per-GPU batch (weak scaling); the head's weight gradients (5 341 556 fp32 =
           21.4 MB, which contain the two DCN weight grads) are all-reduced over NCCL inside every step, started as
           soon as the weight-gradient kernels are done and overlapped with grad_input / grad_offset; the 1/world
           averaging is folded into the weight-gradient kernel's scale.
--impl reference : the CPU deform_conv2d path BASELINE.json names (torchvision.ops.deform_conv2d forward + autograd
           backward, all host threads; the reference itself has no CPU DCN) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]  # P3..P7 of an 800x1344 image
C_IN = C_OUT = 256
BATCH_PER_GPU = 2
EXPLICIT_PREPARE = os.environ.get("SDB_BENCH_PREPARED", "0") != "0"
INDEX_IN_FIRST_HALF = os.environ.get("SDB_BENCH_INDEX_FIRST", "0") != "0"   # measured on 2 GPUs: 0.880 ms vs 0.849 ms for the default
# (the index build in the weight-gradient half delays the start of the all-reduce)
GATHER_OVERLAP = os.environ.get("SDB_BENCH_GATHER_OVERLAP", "0") != "0"   # measured on 2 GPUs: 0.896 ms vs 0.860 ms for the default order
ONE_GRAPH = os.environ.get("SDB_BENCH_ONE_GRAPH", "1") != "0"   # N > 1: capture the all-reduce into the step's CUDA graph
NCCL_CTAS = int(os.environ.get("SDB_BENCH_NCCL_CTAS", "0"))  # developer knob: cap the overlapped all-reduce at this many CTAs and leave
# that many SMs free in the persistent kernels (sdb_set_sm_reserve).  Measured on 2 GPUs: 0 (NCCL default) 0.877 ms, 16 -> 0.877, 8 -> 0.891,
# 4 -> 0.952 (the all-reduce outlasts the data-gradient kernels): SM contention is not what the N > 1 step pays for; default off.
HEAD_PARAMS = 5341556  # RepPoints head parameter count; all-reduced when N > 1
FLOP_PER_PIXEL_PASS = 2 * C_IN * C_OUT * 9  # 1 179 648
METRIC = "reppoints_head_dcn_fwd_bwd_images_per_s"
UNIT = "images/s"


def pixels_per_image():
    return sum(h * w for h, w in LEVELS)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(bf16_tflops=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", 0) or 0),
                    hbm_gbs=float(p["hbm_gbs"]), source="measured")
    except Exception:
        return dict(bf16_tflops=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# reference arm: CPU deform_conv2d path (torchvision), bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, shape, budget_s):
    """Times torchvision.ops.deform_conv2d fwd + autograd bwd on the host (fp32, the only dtype the CPU path supports)
    for ONE of the head's DeformConvs on a [N, 256, H, W] map, `steps` timed runs after `warmup` (stopping early when
    `budget_s` is spent); throughput is scaled by pixel count to whole-step images/s."""
    import torch
    N, H, W = shape
    try:
        from torchvision.ops import deform_conv2d
        kind = "port"  # torchvision's CPU op: same lineage/semantics; the reference itself has no CPU DCN
        what = "torchvision.ops.deform_conv2d"

        def run(x, off, w, gy):
            x.grad = off.grad = w.grad = None
            y = deform_conv2d(x, off, w, None, stride=1, padding=1, dilation=1)
            y.backward(gy)
    except Exception:  # torchvision missing on the box: the repo's C oracle port
        from oracle import dcn as odcn
        kind = "port"
        what = "oracle/dcn_oracle.c (OpenMP)"

        def run(x, off, w, gy):
            odcn.forward(x.detach().numpy(), off.detach().numpy(), w.detach().numpy(), stride=1, padding=1)
            odcn.backward(x.detach().numpy(), off.detach().numpy(), w.detach().numpy(), gy.numpy(), stride=1, padding=1)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, C_IN, H, W, generator=g).requires_grad_()
    w = (torch.randn(C_OUT, C_IN, 3, 3, generator=g) * 0.01).requires_grad_()
    off = (torch.randn(N, 18, H, W, generator=g) * 2.0).requires_grad_()
    gy = torch.randn(N, C_OUT, H, W, generator=g)
    t_start = time.perf_counter()
    for _ in range(warmup):
        run(x, off, w, gy)
        if time.perf_counter() - t_start > budget_s * 0.3:
            break
    t0 = time.perf_counter()
    n = 0
    for _ in range(max(1, steps)):
        run(x, off, w, gy)
        n += 1
        if time.perf_counter() - t_start > budget_s:
            break
    dt = (time.perf_counter() - t0) / n
    sample_px = N * H * W  # one DCN over N*H*W pixels
    step_px = 2 * BATCH_PER_GPU * pixels_per_image()  # two DCNs, batch 2, all levels
    scale = step_px / sample_px
    value = BATCH_PER_GPU / (dt * scale)
    sample = ("%s fwd+bwd (fp32, %d threads) on ONE DeformConv (256->256, 3x3) over a %dx256x%dx%d map, %d timed runs of "
              "%.2f s each; the full step (2 DCN x batch 2 x P3-P7 = %d DCN-pixels) is %.2fx that many pixels, value = "
              "batch / (sample time x %.2f)" % (what, cores, N, H, W, n, dt, step_px, scale, scale))
    return dict(value=value, sample_ms=dt * 1e3, scale=scale, cores=cores, kind=kind, sample=sample, steps=n)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
# offsets of the synthetic workload: dcn_offset ~ N(0, sigma^2) pixels (SURVEY 8d regime ii; 2 px = a trained RepPoints head).
# SDB_BENCH_SIGMA is a developer knob for locality experiments, never set for a reported line.
OFFSET_SIGMA = float(os.environ.get("SDB_BENCH_SIGMA", "2.0"))


class Workload:
    """All device buffers of one step + the C-ABI call sequence (graph-capturable: no allocation, no synchronisation,
    fixed pointers).  Problems are ordered (level, branch); both branches of a level share the level's offsets."""

    def __init__(self, torch, lib_mod, device, seed, batch, levels=LEVELS, modulated=False, scale=1.0, bucket=True,
                 save_columns=os.environ.get("SDB_DCN_SAVE_COLUMNS", "1") != "0"):
        self.torch, self._lib, self.device, self.batch, self.levels_hw = torch, lib_mod, device, batch, levels
        L, lib = lib_mod, lib_mod.lib()
        g = torch.Generator(device="cpu").manual_seed(seed)
        bf = torch.bfloat16
        mk = lambda *s: torch.randn(*s, generator=g)
        self.scale = scale
        self.weights = [(mk(C_OUT, C_IN, 3, 3) * 0.01).to(device, bf) for _ in range(2)]  # cls / refine DCN
        self.biases = [mk(C_OUT).to(device, bf) for _ in range(2)] if modulated else [None, None]
        # the two DCN weight gradients live inside the head's flat gradient bucket (all-reduced when N > 1);
        # the weight-gradient kernel accumulates straight into these views
        from slenderobjdet_b200.dist import GradBucket
        shapes = {"cls_dcn.weight": (C_OUT, C_IN, 3, 3), "refine_dcn.weight": (C_OUT, C_IN, 3, 3)}
        if modulated:
            shapes.update({"cls_dcn.bias": (C_OUT,), "refine_dcn.bias": (C_OUT,)})
        self.bucket = GradBucket(shapes, device, pad_to=HEAD_PARAMS if bucket else 0)
        self.gw = [self.bucket.views["cls_dcn.weight"], self.bucket.views["refine_dcn.weight"]]
        self.gb = [self.bucket.views["cls_dcn.bias"], self.bucket.views["refine_dcn.bias"]] if modulated else [None, None]
        self.geom = L.Geom(batch, C_IN, levels[0][0], levels[0][1], C_OUT, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
        gp = ctypes.byref(self.geom)
        self.io, self.mth = L.SDB_BF16, L.SDB_MATH_BF16
        pbytes = lib.sdb_dcn_prepared_weight_bytes(gp, self.io, self.mth)
        self.prepared = [torch.empty(pbytes, dtype=torch.uint8, device=device) for _ in range(2)]
        self.lv, rows = [], []
        for li, (H, W) in enumerate(levels):
            gl = L.Geom(batch, C_IN, H, W, C_OUT, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
            pkb = lib.sdb_dcn_packed_input_bytes(ctypes.byref(gl), self.mth)
            clb = lib.sdb_dcn_columns_bytes(ctypes.byref(gl), self.mth) if save_columns else 0
            lv = dict(H=H, W=W, off=(mk(batch, 18, H, W) * OFFSET_SIGMA).to(device), br=[],
                      mask=torch.sigmoid(mk(batch, 9, H, W)).to(device) if modulated else None)
            for b in range(2):
                br = dict(x=mk(batch, C_IN, H, W).to(device, bf), gy=mk(batch, C_OUT, H, W).to(device, bf),
                          out=torch.empty(batch, C_OUT, H, W, device=device, dtype=bf),
                          gx=torch.empty(batch, C_IN, H, W, device=device, dtype=bf),
                          goff=torch.empty(batch, 18, H, W, device=device),
                          gmask=torch.empty(batch, 9, H, W, device=device) if modulated else None,
                          pk=torch.empty(max(1, pkb), dtype=torch.uint8, device=device),
                          cols=torch.empty(clb, dtype=torch.uint8, device=device) if clb else None)
                lv["br"].append(br)
                rows.append(L.Problem(batch, H, W, b, li, 0, L.addr(br["x"]), L.addr(lv["off"]), L.addr(lv["mask"]),
                                      L.addr(br["out"]), L.addr(br["pk"]), L.addr(br["gy"]), L.addr(br["gx"]),
                                      L.addr(br["goff"]), L.addr(br["gmask"]), L.addr(br["cols"])))
            self.lv.append(lv)
        self.n = len(rows)
        self.probs = (L.Problem * self.n)(*rows)
        self.wts = (L.Weights * 2)(*[L.Weights(L.addr(self.weights[k]), L.addr(self.biases[k]),
                                               L.addr(self.prepared[k]) if EXPLICIT_PREPARE else None,
                                               L.addr(self.gw[k]), L.addr(self.gb[k])) for k in range(2)])
        wsf = lib.sdb_dcn_multi_workspace_bytes(self.probs, self.n, self.wts, 2, gp, self.io, self.mth, 0)
        wsb = lib.sdb_dcn_multi_workspace_bytes(self.probs, self.n, self.wts, 2, gp, self.io, self.mth, 1)
        self.ws = torch.empty(max(wsf, wsb, 1), dtype=torch.uint8, device=device)
        self.px_step = 2 * batch * sum(h * w for h, w in levels)

    # ---- the three phases of a step, each a handful of launches on `st` -------------------------------------------
    def phase_forward(self, st):
        L, lib, P = self._lib, self._lib.lib(), self._lib.ptr
        sp, gp = ctypes.c_void_p(st.cuda_stream), ctypes.byref(self.geom)
        with self.torch.cuda.stream(st):
            for k in range(2):
                self.gw[k].zero_()
                if self.gb[k] is not None:
                    self.gb[k].zero_()
        # the weights changed (optimiser step): their operand images are rebuilt once per step for all levels -- by the
        # forward / backward calls themselves (prepared == NULL: on the library's side stream, beside the layout packs);
        # SDB_BENCH_PREPARED=1: by explicit sdb_dcn_prepare_weights calls on the step's stream, as before
        for k in range(2 if EXPLICIT_PREPARE else 0):
            L.check(lib.sdb_dcn_prepare_weights(P(self.weights[k]), P(self.biases[k]), gp, self.io, self.mth,
                                                P(self.prepared[k]), sp))
        L.check(lib.sdb_dcn_forward_multi(self.probs, self.n, self.wts, 2, gp, self.io, self.mth, P(self.ws),
                                          self.ws.numel(), sp))

    def phase_backward(self, st, flags=0):
        L, lib, P = self._lib, self._lib.lib(), self._lib.ptr
        L.check(lib.sdb_dcn_backward_multi(self.probs, self.n, self.wts, 2, ctypes.byref(self.geom), self.io, self.mth,
                                           float(self.scale), int(flags), P(self.ws), self.ws.numel(),
                                           ctypes.c_void_p(st.cuda_stream)))

    def step(self, st):
        self.phase_forward(st)
        self.phase_backward(st, 0)


def parity_check(torch, L, device):
    """One level of the benchmarked workload (P5, both branches, batch 2, bf16) through the SAME multi-problem entry
    points, checked against the CPU oracle outside the timed region: worst relative L2 error over out / grad_x /
    grad_offset / grad_weight, against the 1e-2 bf16 tolerance of BASELINE.json."""
    import numpy as np
    from oracle import dcn as odcn
    wl = Workload(torch, L, device, seed=4242, batch=2, levels=[LEVELS[2]], bucket=False)
    st = torch.cuda.current_stream(device)
    wl.step(st)
    torch.cuda.synchronize(device)
    f = lambda t: t.detach().float().cpu().numpy()
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))
    worst, lv = {}, wl.lv[0]
    for b in range(2):
        br = lv["br"][b]
        x, w, off, gy = f(br["x"]), f(wl.weights[b]), f(lv["off"]), f(br["gy"])
        yo = odcn.forward(x, off, w, stride=1, padding=1)
        go = odcn.backward(x, off, w, gy, stride=1, padding=1)
        for k, got, ref in (("out", br["out"], yo), ("grad_x", br["gx"], go["grad_x"]),
                            ("grad_offset", br["goff"], go["grad_offset"]), ("grad_weight", wl.gw[b], go["grad_weight"])):
            worst[k] = max(worst.get(k, 0.0), rel(f(got), ref))
    ok = all(v < 1e-2 for v in worst.values())
    return {"checked": "P5 (25x42) x both branches x batch 2 vs oracle/dcn_oracle.c, relative L2", "tol": 1e-2,
            "rel_err": {k: float("%.3g" % v) for k, v in worst.items()}, "ok": bool(ok)}


def time_workload(torch, wl, stream, steps, warmup, l2_flush, world=1, dist=None, graph_ok=True):
    """-> (ms per step, description of the launch mode).  N > 1: weight gradients first, their all-reduce overlapped
    with the data-gradient kernels."""
    L = wl._lib
    graphs = None
    # N > 1, default order: weight gradients first, then the all-reduce beside grad_offset + grad_input.
    # SDB_BENCH_GATHER_OVERLAP=1 (experiment, slower): grad_offset / grad_mask first, then the weight gradients, then the
    # all-reduce beside the grad_input gather only (an ordinary grid) -- the collective then starts later than it can
    # finish behind a 150 us HBM-latency-bound kernel
    if GATHER_OVERLAP:
        FIRST_HALF = L.SDB_BWD_DATA_ONLY | L.SDB_BWD_NO_GATHER
        SECOND_HALF = L.SDB_BWD_DATA_ONLY | L.SDB_BWD_GATHER_ONLY | L.SDB_BWD_GRAD_PACKED
    elif INDEX_IN_FIRST_HALF:   # the transposed index is built beside the weight-gradient GEMM of the first half
        FIRST_HALF = L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_BUILD_INDEX
        SECOND_HALF = L.SDB_BWD_DATA_ONLY | L.SDB_BWD_GRAD_PACKED | L.SDB_BWD_INDEX_READY
    else:
        FIRST_HALF = L.SDB_BWD_WEIGHT_ONLY
        SECOND_HALF = L.SDB_BWD_DATA_ONLY | L.SDB_BWD_GRAD_PACKED
    with torch.cuda.stream(stream):
        wl.step(stream)
        stream.synchronize()
        def capture_single():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                wl.step(stream)
            return [g]

        def capture_with_allreduce():
            # the whole step INCLUDING the all-reduce as one graph: the collective is captured on NCCL's stream between
            # the weight-gradient kernels and the join, beside the data-gradient kernels -- no host launch gaps
            for _ in range(2):   # NCCL warm-up outside the capture (communicator setup, buffer registration)
                w0 = wl.bucket.all_reduce(average=True, async_op=True, prescaled=True)
                if w0 is not None:
                    w0.wait()
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                wl.phase_forward(stream)
                wl.phase_backward(stream, FIRST_HALF)
                if GATHER_OVERLAP:
                    wl.phase_backward(stream, L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_GRAD_PACKED)
                work = wl.bucket.all_reduce(average=True, async_op=True, prescaled=True)
                wl.phase_backward(stream, SECOND_HALF)
                if work is not None:
                    work.wait()
            return [g]

        def capture_two_halves():
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, stream=stream):
                wl.phase_forward(stream)
                wl.phase_backward(stream, FIRST_HALF)
                if GATHER_OVERLAP:
                    wl.phase_backward(stream, L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_GRAD_PACKED)
            L.lib().sdb_set_sm_reserve(max(NCCL_CTAS, 0))   # grids are baked into the graph at capture
            with torch.cuda.graph(gb, stream=stream):
                wl.phase_backward(stream, SECOND_HALF)
            L.lib().sdb_set_sm_reserve(0)
            return [ga, gb]

        if graph_ok:
            plans = [capture_single] if world == 1 else ([capture_with_allreduce] if ONE_GRAPH else []) + [capture_two_halves]
            for plan in plans:
                try:
                    graphs = plan()
                    break
                except Exception as e:  # next plan, or eager launches: say so
                    graphs = None
                    sys.stderr.write("bench.py: CUDA graph capture (%s) failed (%r)\n" % (plan.__name__, e))
                    torch.cuda.synchronize()

    def one_step():
        if world == 1:
            if graphs:
                graphs[0].replay()
            else:
                wl.step(stream)
            return
        if graphs and len(graphs) == 1:   # step and collective in one graph
            graphs[0].replay()
            return
        if graphs:
            graphs[0].replay()
        else:
            wl.phase_forward(stream)
            wl.phase_backward(stream, FIRST_HALF)
            if GATHER_OVERLAP:
                wl.phase_backward(stream, L.SDB_BWD_WEIGHT_ONLY | L.SDB_BWD_GRAD_PACKED)
        # the weight gradients (already scaled by 1/world in the kernel) are complete: reduce them on NCCL's stream
        # while grad_input / grad_offset run
        work = wl.bucket.all_reduce(average=True, async_op=True, prescaled=True)
        if graphs:
            graphs[1].replay()
        else:
            L.lib().sdb_set_sm_reserve(max(NCCL_CTAS, 0))
            wl.phase_backward(stream, SECOND_HALF)
            L.lib().sdb_set_sm_reserve(0)
        if work is not None:
            work.wait()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            one_step()
        sync_all()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            l2_flush.zero_()          # flush L2 between timed iterations (outside the event pair)
            ev[i][0].record(stream)
            one_step()
            ev[i][1].record(stream)
        sync_all()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    mode = ("cuda_graph" if graphs else "eager") + ", one launch per kernel over all levels and both branches"
    if world > 1 and graphs:
        mode += ", all-reduce captured in the step's graph" if len(graphs) == 1 else ", two graphs around an eager all-reduce"
    return ms, mode


def extra_paths(torch, sdb, device, l2_flush):
    """The round's other tensor-core paths through the operator API, as CUDA-graph replays (median of 10, L2 flushed):
    the forward of float32 tensors (SIMT fp32 vs kind::tf32, one DCN on the P3 map) and one tower layer
    (conv 3x3 -> GroupNorm(32) -> ReLU, reppointsv2.py:644-675) over P3-P7 x 2 towers, forward + backward."""
    import slenderobjdet_b200.layers as LL
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(0)

    def replay_us(fn, iters=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            l2_flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return round(ts[len(ts) // 2], 1)

    out = {}
    H, W = LEVELS[0]
    x = torch.randn(BATCH_PER_GPU, C_IN, H, W, generator=g).to(device)
    w = (torch.randn(C_OUT, C_IN, 3, 3, generator=g) * 0.01).to(device)
    off = (torch.randn(BATCH_PER_GPU, 18, H, W, generator=g) * 2).to(device)
    f32 = {}
    ref = None
    for mode in ("fp32", "tf32x3", "tf32"):
        def fwd():
            with sdb.dcn_math(mode), torch.no_grad():
                return sdb.deform_conv(x, off, w, 1, 1, 1, 1, 1)
        f32[mode + "_us"] = replay_us(fwd)
        y = fwd().double()
        ref = y if ref is None else ref
        f32[mode + "_rel_err_vs_fp32"] = float("%.2e" % float((y - ref).norm() / ref.norm()))
    out["float32 tensors: forward of one DeformConv 256->256 on the P3 map, batch 2 (fp32 = SIMT kernel; tf32x3 is what 'auto' runs)"] = f32
    xs = [torch.randn(BATCH_PER_GPU, C_IN, h, w_, generator=g).to(bf).to(device).requires_grad_() for _ in range(2) for (h, w_) in LEVELS]
    gys = [torch.randn(BATCH_PER_GPU, C_IN, h, w_, generator=g).to(bf).to(device) for _ in range(2) for (h, w_) in LEVELS]
    ids = [t for t in range(2) for _ in LEVELS]
    ws = [(torch.randn(C_OUT, C_IN, 3, 3, generator=g) * 0.01).to(bf).to(device).requires_grad_() for _ in range(2)]
    gam = [torch.ones(C_OUT, device=device).requires_grad_() for _ in range(2)]
    bet = [torch.zeros(C_OUT, device=device).requires_grad_() for _ in range(2)]

    def conv_fb():
        torch.autograd.backward(LL.conv2d_multi(xs, ws, None, 1, 1, ids), gys)

    def gn_fb():
        torch.autograd.backward(LL.group_norm_relu_multi(xs, gam, bet, 32, 1e-5, ids), gys)
    px = sum(h * w_ for (h, w_) in LEVELS) * BATCH_PER_GPU * 2
    tc, tg = replay_us(conv_fb), replay_us(gn_fb)
    out["tower layer (conv 3x3 -> GroupNorm(32) -> ReLU) over P3-P7 x 2 towers, batch 2, bf16, forward + backward"] = {
        "conv_us": tc, "conv_tflops_per_s": round(3 * 2.0 * px * C_IN * C_OUT * 9 / tc / 1e6, 1),
        "groupnorm_relu_us": tg, "groupnorm_relu_gb_per_s": round(8 * px * C_OUT * 2 / tg / 1e3, 1)}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from slenderobjdet_b200 import _lib as L
    import slenderobjdet_b200 as sdb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # the weight-gradient all-reduce runs BESIDE the persistent data-gradient kernels: cap NCCL at NCCL_CTAS CTAs and
        # leave that many SMs free (sdb_set_sm_reserve) so neither waits for the other's SMs
        if NCCL_CTAS > 0:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = NCCL_CTAS
            opts.config.min_ctas = 1
            dist.init_process_group("nccl", device_id=device, pg_options=opts)
        else:   # developer knob: NCCL's own CTA count, nothing reserved
            dist.init_process_group("nccl", device_id=device)
    lib = L.lib()
    if os.environ.get("SDB_BENCH_PAIR_BWD") is not None:   # developer knob: CTA-pair grad_offset kernel on / off
        lib.sdb_set_backward_pair(int(os.environ["SDB_BENCH_PAIR_BWD"]))
    peaks = load_peaks()
    batch = BATCH_PER_GPU
    stream = torch.cuda.Stream(device)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    parity = parity_check(torch, L, device) if rank == 0 else None
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: parity check against the oracle FAILED: %r" % (parity,))

    wl = Workload(torch, L, device, seed=rank, batch=batch, scale=1.0 / world)
    px_step = wl.px_step                                  # DCN-pixels per step per GPU
    flops_step = 3 * FLOP_PER_PIXEL_PASS * px_step        # fwd + dgrad + wgrad (algorithmic; grad_input is a 4th executed pass)

    with torch.cuda.stream(stream):
        n0 = lib.sdb_launch_count()
        wl.step(stream)
        launches_per_step = lib.sdb_launch_count() - n0
        torch_fills_per_step = 2                          # zero_() of the two weight-gradient views
        stream.synchronize()

    # ---- `value`: device-resident, timed with CUDA events on the launching stream --------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, mode = time_workload(torch, wl, stream, args.steps, args.warmup, l2_flush, world, dist if world > 1 else None,
                                  graph_ok=not args.no_graph)
    t = torch.tensor([step_ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = world * batch / (step_ms * 1e-3)

    # ---- roofline: per-kernel durations from the library's own event pairs ----------------------
    reps = 5
    with torch.cuda.stream(stream):
        lib.sdb_profile_reset()
        lib.sdb_profile_enable(1)
        for _ in range(reps):
            l2_flush.zero_()
            wl.step(stream)              # one stream: the event pairs time each tensor-core kernel alone
        lib.sdb_profile_enable(0)
        stream.synchronize()
    names = ["dcn_fwd_tc_kernel", "dcn_bwd_data_tc_kernel (grad_offset)", "dcn_wgrad_col_tc_kernel",
             "dcn_dx_gather_kernel (grad_input)"]
    # the first three are GEMM passes (tensor-bound on paper); grad_input is a gather of the exported dcol tiles over the
    # transposed index (HBM-bound on paper): dcol read once (taps * C * 2 B per output pixel) + the index (36 entries of
    # 8 B per input pixel, shared by the two branches of a level) + grad_x written (C * 2 B per input pixel)
    gather_bytes = px_step * (9 * C_IN * 2 + C_IN * 2) + (px_step // 2) * 36 * 8
    kern = []
    for slot in range(4):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        L.check(lib.sdb_profile_read(slot, ctypes.byref(ms), ctypes.byref(n)))
        kern.append((ms.value / reps, n.value // reps))
    dom = max(range(3), key=lambda s: kern[s][0])   # dominant GEMM kernel
    dom_ms, dom_launches = kern[dom]
    # algorithmic FLOPs of one GEMM pass over every (level, branch) = FLOP_PER_PIXEL_PASS * pixels
    achieved = FLOP_PER_PIXEL_PASS * px_step / (dom_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": names[dom], "achieved": round(achieved, 2),
                "peak": peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst bf16, cuBLAS)",
                "unit": "TFLOP/s", "frac": round(achieved / peaks["bf16_tflops"], 4),
                "launches_per_step": dom_launches, "avg_launch_us": round(dom_ms * 1e3 / max(dom_launches, 1), 2),
                "algorithmic_flops_per_launch": FLOP_PER_PIXEL_PASS * px_step,
                "algorithmic_flops_per_step": flops_step, "executed_gemm_flops_per_step": 3 * FLOP_PER_PIXEL_PASS * px_step,
                "per_kernel_ms_per_step": {names[s]: round(kern[s][0], 4) for s in range(4)},
                "per_kernel_frac_of_peak": {names[s]: round(FLOP_PER_PIXEL_PASS * px_step / (kern[s][0] * 1e-3) / 1e12
                                                            / peaks["bf16_tflops"], 4) for s in range(3) if kern[s][0] > 0},
                "grad_input_gather": ({"bound": "hbm", "algorithmic_bytes_per_launch": gather_bytes,
                                       "achieved": round(gather_bytes / (kern[3][0] * 1e-3) / 1e9, 1),
                                       "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                       "frac": round(gather_bytes / (kern[3][0] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)
                                       if peaks.get("hbm_gbs") else None} if kern[3][0] > 0 else None),
                "step_frac_of_peak": round(flops_step / (step_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], 4),
                "traffic": None}
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get(names[dom])
        except Exception:
            pass

    # ---- the other single-GPU configurations of BASELINE.json, as extra fields (rank 0, N = 1 only) --------------
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        for tag, kw in (("configs[2] modulated DCNv2 (mask + bias), batch 8", dict(batch=8, modulated=True)),
                        ("configs[4] per-GPU batch 16", dict(batch=16))):
            try:
                w2 = Workload(torch, L, device, seed=7, **kw)
                ms2, _ = time_workload(torch, w2, stream, 5, 3, l2_flush)
                fl2 = 3 * FLOP_PER_PIXEL_PASS * w2.px_step
                extra[tag] = {"ms_per_step": round(ms2, 4), "images_per_s": round(kw["batch"] / (ms2 * 1e-3), 1),
                              "tflops_per_s": round(fl2 / (ms2 * 1e-3) / 1e12, 1),
                              "frac_of_peak": round(fl2 / (ms2 * 1e-3) / 1e12 / peaks["bf16_tflops"], 4)}
                del w2
                torch.cuda.empty_cache()
            except Exception as e:  # never lose the headline line to an extra
                extra[tag] = {"error": repr(e)[:200]}

        try:
            extra.update(extra_paths(torch, sdb, device, l2_flush))
        except Exception as e:
            extra["float32 forward / tower layer"] = {"error": repr(e)[:200]}

    # ---- e2e: public Python API, host buffers, H2D + D2H inside the timed region -----------------
    e2e = measure_e2e(torch, sdb, device, stream, batch, args, world, dist if world > 1 else None, rank)

    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()

    line = None
    if rank == 0:
        if args.no_cpu:   # developer flag (A/B timing of two builds): never used for a reported line
            cb = dict(value=0.0, cores=0, kind="skipped", sample="skipped (--no-cpu developer flag)")
        else:
            cb = cpu_reference(steps=1, warmup=0, shape=(2, 100, 152), budget_s=60.0)   # BASELINE.json configs[0] shape
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "RepPoints-v1 R-50-FPN head DCN pair, P3-P7 @800x1344, batch %d/GPU, fwd+bwd "
                                   "(BASELINE.json configs[1])" % batch,
                       "global_batch": world * batch, "levels": LEVELS, "channels": [C_IN, C_OUT],
                       "parallelism": "dp%d" % world, "l2": "flushed between timed iterations (256 MiB memset)",
                       "launch": mode, "launches_per_step": int(launches_per_step + torch_fills_per_step),
                       "tflops_per_s": round(flops_step * world / (step_ms * 1e-3) / 1e12, 2),
                       "allreduce_bytes": HEAD_PARAMS * 4 if world > 1 else 0,
                       "allreduce": ("overlapped with the grad_input gather (after grad_offset and the weight gradients), 1/world folded into the kernel" if GATHER_OVERLAP else "overlapped with grad_offset + grad_input, 1/world folded into the kernel") if world > 1 else None},
            "roofline": roofline,
            "parity": parity,
            "cpu_baseline": {"value": round(cb["value"], 5), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                             "sample": cb["sample"]},
            "e2e": e2e,
            "other_configs": extra,
            "gpu_launches": int((launches_per_step + torch_fills_per_step) * args.steps),
            "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def measure_e2e(torch, sdb, device, stream, batch, args, world, dist, rank):
    """The step through the public operator API, inputs in pinned host memory.  Two device buffer sets: the copy stream
    uploads step i+1 while the compute stream runs step i (every step's upload and download are inside the timed region)."""
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(100 + rank)
    convs = [sdb.DeformConv(C_IN, C_OUT, 3, 1, 1).to(device, bf) for _ in range(2)]
    ws = [c.weight for c in convs]
    host = []
    for (H, W) in LEVELS:
        host.append(dict(
            x=[torch.randn(batch, C_IN, H, W, generator=g).to(bf).pin_memory() for _ in range(2)],
            gy=[torch.randn(batch, C_OUT, H, W, generator=g).to(bf).pin_memory() for _ in range(2)],
            off=(torch.randn(batch, 18, H, W, generator=g) * 2.0).pin_memory()))
    h2d = sum(sum(t.numel() * t.element_size() for t in lv["x"] + lv["gy"]) + lv["off"].numel() * 4 for lv in host)
    out_host = [torch.empty(C_OUT, C_IN, 3, 3, dtype=bf).pin_memory() for _ in range(2)]
    goff_host = [torch.empty(batch, 18, H, W).pin_memory() for (H, W) in LEVELS]
    d2h = sum(t.numel() * t.element_size() for t in out_host) + sum(t.numel() * 4 for t in goff_host)
    from slenderobjdet_b200.dist import GradBucket
    bucket = GradBucket({"cls_dcn.weight": (C_OUT, C_IN, 3, 3), "refine_dcn.weight": (C_OUT, C_IN, 3, 3)},
                        device, pad_to=HEAD_PARAMS) if world > 1 else None
    copy_stream = torch.cuda.Stream(device)
    sets = []
    for _ in range(2):
        sets.append(dict(x=[[torch.empty_like(t, device=device) for t in lv["x"]] for lv in host],
                         gy=[[torch.empty_like(t, device=device) for t in lv["gy"]] for lv in host],
                         off=[torch.empty_like(lv["off"], device=device) for lv in host],
                         ready=torch.cuda.Event(), free=torch.cuda.Event()))
    wids = [b for _ in LEVELS for b in range(2)]

    def upload(s):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(s["free"])       # the step that last read this set has finished
            for li, lv in enumerate(host):
                s["off"][li].copy_(lv["off"], non_blocking=True)
                for b in range(2):
                    s["x"][li][b].copy_(lv["x"][b], non_blocking=True)
                    s["gy"][li][b].copy_(lv["gy"][b], non_blocking=True)
            s["ready"].record(copy_stream)

    def compute(s):
        main = torch.cuda.current_stream()
        main.wait_event(s["ready"])
        for w in ws:
            w.grad = None
        offs_leaf = [o.detach().requires_grad_() for o in s["off"]]
        xs = [s["x"][li][b].detach().requires_grad_() for li in range(len(LEVELS)) for b in range(2)]
        offs = [offs_leaf[li] for li in range(len(LEVELS)) for _ in range(2)]   # both branches: the level's one offset
        gys = [s["gy"][li][b] for li in range(len(LEVELS)) for b in range(2)]
        ys = sdb.deform_conv_multi(xs, offs, ws, 1, 1, 1, weight_ids=wids)
        torch.autograd.backward(ys, gys)
        s["free"].record(main)
        if world > 1:
            params = {"cls_dcn.weight": convs[0].weight, "refine_dcn.weight": convs[1].weight}
            bucket.pack({k: p.grad for k, p in params.items()})
            bucket.all_reduce(average=True)
            bucket.unpack(params)
        for b in range(2):
            out_host[b].copy_(convs[b].weight.grad, non_blocking=True)
        for i, off in enumerate(offs_leaf):
            goff_host[i].copy_(off.grad, non_blocking=True)

    def run(nsteps):
        main = torch.cuda.current_stream()
        upload(sets[0])
        for i in range(nsteps):
            if i + 1 < nsteps:
                upload(sets[(i + 1) % 2])
            compute(sets[i % 2])
            main.synchronize()   # the step's results are on the host

    with torch.cuda.stream(stream):
        for s in sets:
            s["free"].record(stream)
        run(max(3, args.warmup))
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(args.steps)
        e1.record(stream)
        torch.cuda.synchronize(device)
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        ms = max(e0.elapsed_time(e1) / args.steps, wall_ms)  # host-side work is part of an end-to-end step
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": round(world * batch / (ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(ms, 4),
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "api": "slenderobjdet_b200.deform_conv_multi (10 problems, 2 weights) + autograd backward; pinned host tensors; "
                   "uploads of step i+1 overlap the kernels of step i (two device buffer sets)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # a step of this arm = the P4 map (2 x 256 x 50 x 84) through one DeformConv, fwd+bwd; the whole run is bounded
    cb = cpu_reference(steps=args.steps, warmup=min(args.warmup, 2), shape=(2, 50, 84), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(cb["value"], 5), "unit": UNIT, "n_gpus": world,
        "steps": cb["steps"], "warmup": min(args.warmup, 2), "ms_per_step": round(cb["sample_ms"], 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RepPoints-v1 R-50-FPN head DCN pair, P3-P7 @800x1344, batch %d, fwd+bwd "
                               "(BASELINE.json configs[1]); CPU deform_conv2d path, each step = a bounded sample "
                               "(ms_per_step is the sample's time; value is scaled to the whole step by pixels: x%.2f)"
                               % (BATCH_PER_GPU, cb["scale"]),
                   "global_batch": BATCH_PER_GPU, "parallelism": "cpu"},
        "cpu_baseline": {"value": round(cb["value"], 5), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                         "sample": cb["sample"]},
        "e2e": {"value": round(cb["value"], 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager C-ABI launches instead of a CUDA graph")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[4] extra measurements")
    ap.add_argument("--no-cpu", action="store_true", help="developer flag: skip the CPU baseline sample (A/B timing of builds)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun on this node
        port = os.environ.get("MASTER_PORT", "29531")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--no-graph"] if args.no_graph else [])
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
