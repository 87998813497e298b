#!/usr/bin/env python
"""bench.py -- RepPoints-head DCN forward+backward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): RepPoints-v1 R-50-FPN head, 800x1333 input padded to 800x1344,
FPN levels P3-P7 (100x168 ... 7x11, 22 400 points / image), batch 2 per GPU, the head's two
DeformConv(256,256,3,1,1) (cls + pts-refine branch, weights shared across levels) applied to
every level, forward + backward (grad_input, grad_offset, grad_weight), bf16 tensors, fp32
accumulate.  One "step" = that whole pass (20 deformable convolutions fwd+bwd).  Synthetic data
(SURVEY.md 8d): randn features, weights randn*0.01, offsets randn*2 px.

`value`  : images/s with every input resident in HBM, the step replayed as ONE CUDA graph of the
           library's C-ABI calls (timed with CUDA events on the launching stream, max over ranks).
`e2e`    : images/s through the public Python operator API (slenderobjdet_b200.DeformConv +
           autograd) with all inputs in pinned HOST memory: H2D copies of features / offsets /
           grad_out and the D2H read of the gradients are inside the timed region.
N > 1    : one process per GPU, same per-GPU batch (weak scaling); the head's weight gradients
           (5 341 556 fp32 = 21.4 MB, which contain the two DCN weight grads) are all-reduced over
           NCCL inside every step.
--impl reference : the CPU deform_conv2d path BASELINE.json names (torchvision.ops.deform_conv2d
           forward + autograd backward, all host threads) on a bounded sample of the same workload.
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
        return dict(bf16_tflops=float(p["bf16_tflops"]), hbm_gbs=float(p["hbm_gbs"]), source="measured")
    except Exception:
        return dict(bf16_tflops=1590.0, hbm_gbs=6650.0, source="fallback")


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
def cpu_reference(steps, warmup, seconds_hint=8.0):
    """Times torchvision.ops.deform_conv2d fwd + autograd bwd on the host.  The sample is one P5
    level (25x42) of ONE image through ONE of the head's DeformConvs (fp32, the only dtype the CPU
    path supports); throughput is scaled by pixel count to whole-step images/s."""
    import torch
    try:
        from torchvision.ops import deform_conv2d
        kind = "port"  # torchvision's CPU op: same lineage/semantics; the reference itself has no CPU DCN
        def run(x, off, w, gy):
            x.grad = off.grad = w.grad = None
            y = deform_conv2d(x, off, w, None, stride=1, padding=1, dilation=1)
            y.backward(gy)
    except Exception:  # torchvision missing on the box: the repo's C oracle port
        from oracle import dcn as odcn
        kind = "port"
        def run(x, off, w, gy):
            odcn.forward(x.detach().numpy(), off.detach().numpy(), w.detach().numpy(), stride=1, padding=1)
            odcn.backward(x.detach().numpy(), off.detach().numpy(), w.detach().numpy(), gy.numpy(), stride=1, padding=1)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, W = LEVELS[2]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, C_IN, H, W, generator=g).requires_grad_()
    w = (torch.randn(C_OUT, C_IN, 3, 3, generator=g) * 0.01).requires_grad_()
    off = (torch.randn(1, 18, H, W, generator=g) * 2.0).requires_grad_()
    gy = torch.randn(1, C_OUT, H, W, generator=g)
    for _ in range(max(1, min(warmup, 2))):
        run(x, off, w, gy)
    t0 = time.perf_counter()
    n = 0
    for _ in range(max(1, steps)):
        run(x, off, w, gy)
        n += 1
        if time.perf_counter() - t0 > seconds_hint * 3:
            break
    dt = (time.perf_counter() - t0) / n
    sample_px = H * W  # one DCN over H*W pixels
    step_px = 2 * BATCH_PER_GPU * pixels_per_image()  # two DCNs, batch 2, all levels
    ms_per_step = dt * 1e3 * step_px / sample_px
    value = BATCH_PER_GPU / (ms_per_step * 1e-3)
    sample = ("torchvision.ops.deform_conv2d fwd+bwd (fp32, %d threads) on 1 image x P5 (25x42) x 1 DeformConv "
              "(256->256, 3x3), %d timed runs of %.2f s; scaled by pixels to the full step (2 DCN x batch 2 x P3-P7)"
              % (cores, n, dt))
    return dict(value=value, ms_per_step=ms_per_step, cores=cores, kind=kind, sample=sample, steps=n)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class Workload:
    """All device buffers of one step + the C-ABI call sequence (graph-capturable: no allocation,
    no synchronisation, fixed pointers)."""

    def __init__(self, torch, lib_mod, device, seed, batch):
        self.torch, self._lib, self.device, self.batch = torch, lib_mod, device, batch
        lib = lib_mod.lib()
        g = torch.Generator(device="cpu").manual_seed(seed)
        bf = torch.bfloat16
        mk = lambda *s: torch.randn(*s, generator=g)
        self.weights = [(mk(C_OUT, C_IN, 3, 3) * 0.01).to(device, bf) for _ in range(2)]  # cls / refine DCN
        # the two DCN weight gradients live inside the head's flat gradient bucket (all-reduced when N > 1);
        # sdb_dcn_backward_weight accumulates straight into these views
        from slenderobjdet_b200.dist import GradBucket
        self.bucket = GradBucket({"cls_dcn.weight": (C_OUT, C_IN, 3, 3), "refine_dcn.weight": (C_OUT, C_IN, 3, 3)},
                                 device, pad_to=HEAD_PARAMS)
        self.gw = [self.bucket.views["cls_dcn.weight"], self.bucket.views["refine_dcn.weight"]]
        self.levels = []
        for (H, W) in LEVELS:
            geom = lib_mod.Geom(batch, C_IN, H, W, C_OUT, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
            gp = ctypes.byref(geom)
            lv = dict(geom=geom, H=H, W=W, off=(mk(batch, 18, H, W) * 2.0).to(device), br=[])
            lv["goff"] = [torch.empty_like(lv["off"]) for _ in range(2)]
            wsb = [lib.sdb_dcn_workspace_bytes(op, gp, lib_mod.SDB_BF16, lib_mod.SDB_MATH_BF16) for op in range(3)]
            pkb = lib.sdb_dcn_packed_input_bytes(gp, lib_mod.SDB_MATH_BF16)
            for b in range(2):
                lv["br"].append(dict(
                    x=mk(batch, C_IN, H, W).to(device, bf), gy=mk(batch, C_OUT, H, W).to(device, bf),
                    out=torch.empty(batch, C_OUT, H, W, device=device, dtype=bf),
                    gx=torch.zeros(batch, C_IN, H, W, device=device, dtype=bf),
                    ws=[torch.empty(max(1, n), dtype=torch.uint8, device=device) for n in wsb],
                    pk=torch.empty(max(1, pkb), dtype=torch.uint8, device=device)))
            self.levels.append(lv)

    def step(self, main, serial=False):
        """forward + backward_data + backward_weight of both DCNs on every level, via the C ABI.
        The ten (level, branch) problems are independent, so each runs on its own stream, forked
        from / joined to `main` with events (under CUDA-graph capture these become graph edges):
        the P5-P7 maps are one to nine 128-pixel tiles and would otherwise leave 140+ SMs idle.
        The weight gradient of a branch accumulates over levels, so its five calls stay in order on
        one stream per branch."""
        torch = self.torch
        L, lib, P = self._lib, self._lib.lib(), self._lib.ptr
        io, mth = L.SDB_BF16, L.SDB_MATH_BF16
        if not hasattr(self, "side_streams"):
            self.side_streams = [torch.cuda.Stream(self.device) for _ in range(2 * len(self.levels) + 2)]
        # serial=True (per-kernel profiling pass): everything in order on `main`
        self.side = [main] * len(self.side_streams) if serial else self.side_streams
        sp = lambda st: ctypes.c_void_p(st.cuda_stream)
        with torch.cuda.stream(main):
            for gw in self.gw:
                gw.zero_()
        # ---- forward of every (level, branch) ----
        k = 0
        for lv in self.levels:
            gp = ctypes.byref(lv["geom"])
            for b, br in enumerate(lv["br"]):
                st = self.side[k]; k += 1
                st.wait_stream(main)
                L.check(lib.sdb_dcn_forward(P(br["x"]), P(lv["off"]), None, P(self.weights[b]), None, P(br["out"]), gp,
                                            io, mth, P(br["ws"][0]), br["ws"][0].numel(), P(br["pk"]), sp(st)))
        for st in self.side[:k]:
            main.wait_stream(st)
        # ---- backward ----
        for st in self.side:
            st.wait_stream(main)
        k = 0
        nl = 2 * len(self.levels)
        for lv in reversed(self.levels):
            gp = ctypes.byref(lv["geom"])
            for b, br in enumerate(lv["br"]):
                st = self.side[k]; k += 1
                with torch.cuda.stream(st):
                    br["gx"].zero_()
                L.check(lib.sdb_dcn_backward_data(P(br["x"]), P(lv["off"]), None, P(self.weights[b]), P(br["gy"]),
                                                  P(br["gx"]), P(lv["goff"][b]), None, gp, io, mth, P(br["ws"][1]),
                                                  br["ws"][1].numel(), P(br["pk"]), sp(st)))
                L.check(lib.sdb_dcn_backward_weight(P(br["x"]), P(lv["off"]), None, P(br["gy"]), P(self.gw[b]), None,
                                                    1.0, gp, io, mth, P(br["ws"][2]), br["ws"][2].numel(), P(br["pk"]),
                                                    sp(self.side[nl + b])))
        for st in self.side:
            main.wait_stream(st)


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
        dist.init_process_group("nccl", device_id=device)
    lib = L.lib()
    peaks = load_peaks()
    batch = BATCH_PER_GPU
    wl = Workload(torch, L, device, seed=rank, batch=batch)
    stream = torch.cuda.Stream(device)
    px_step = 2 * batch * pixels_per_image()              # DCN-pixels per step per GPU
    flops_step = 3 * FLOP_PER_PIXEL_PASS * px_step        # fwd + dgrad + wgrad

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    # ---- launch count + CUDA graph of one step ------------------------------------------------
    with torch.cuda.stream(stream):
        n0 = lib.sdb_launch_count()
        wl.step(stream)
        launches_per_step = lib.sdb_launch_count() - n0
        # torch-side zero_() fills inside the step: 2 (gw) + 10 (gx)
        torch_fills_per_step = 2 + 2 * len(LEVELS)
        stream.synchronize()
        graph = None
        if not args.no_graph:
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    wl.step(stream)
            except Exception as e:  # keep running eagerly, say so in config
                graph = None
                sys.stderr.write("bench.py: CUDA graph capture failed (%r); timing eager launches\n" % (e,))
                torch.cuda.synchronize(device)

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            wl.step(stream)
        if world > 1:
            wl.bucket.all_reduce(average=True)   # one NCCL all-reduce of the 21.4 MB head-gradient bucket

    # ---- `value`: device-resident, timed with CUDA events on the launching stream --------------
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            one_step()
        sync_all()
        sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            l2_flush.zero_()          # flush L2 between timed iterations (outside the event pair)
            ev[i][0].record(stream)
            one_step()
            ev[i][1].record(stream)
        sync_all()
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([step_ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = world * batch / (step_ms * 1e-3)

    # ---- roofline: per-kernel durations from the library's own event pairs ----------------------
    with torch.cuda.stream(stream):
        lib.sdb_profile_reset()
        lib.sdb_profile_enable(1)
        for _ in range(3):
            l2_flush.zero_()
            wl.step(stream, serial=True)   # kernels back to back on one stream: event pairs time each alone
        lib.sdb_profile_enable(0)
        stream.synchronize()
    names = ["dcn_fwd_tc_kernel<MODE_FWD>", "dcn_bwd_data_tc_kernel (grad_offset)", "dcn_bwd_weight_tc_kernel",
             "dcn_fwd_tc_kernel<MODE_DX> (grad_input)"]
    kern = []
    for slot in range(4):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        L.check(lib.sdb_profile_read(slot, ctypes.byref(ms), ctypes.byref(n)))
        kern.append((ms.value / 3.0, n.value // 3))
    dom = max(range(4), key=lambda s: kern[s][0])
    dom_ms, dom_launches = kern[dom]
    # algorithmic FLOPs of one pass over every (level, branch) = FLOP_PER_PIXEL_PASS * pixels
    achieved = FLOP_PER_PIXEL_PASS * px_step / (dom_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": names[dom], "achieved": round(achieved, 2),
                "peak": peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst bf16, cuBLAS)",
                "unit": "TFLOP/s", "frac": round(achieved / peaks["bf16_tflops"], 4),
                "launches_per_step": dom_launches, "avg_launch_us": round(dom_ms * 1e3 / max(dom_launches, 1), 2),
                "algorithmic_flops_per_step": FLOP_PER_PIXEL_PASS * px_step,
                "per_kernel_ms_per_step": {names[s]: round(kern[s][0], 4) for s in range(4)},
                "traffic": None}
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get(names[dom])
        except Exception:
            pass

    # ---- e2e: public Python API, host buffers, H2D + D2H inside the timed region -----------------
    e2e = measure_e2e(torch, sdb, device, stream, batch, args, world, dist if world > 1 else None, rank)

    clocks = None
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()

    line = None
    if rank == 0:
        cb = cpu_reference(steps=3, warmup=1, seconds_hint=5.0)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "RepPoints-v1 R-50-FPN head DCN pair, P3-P7 @800x1344, batch %d/GPU, fwd+bwd "
                                   "(BASELINE.json configs[1])" % batch,
                       "global_batch": world * batch, "levels": LEVELS, "channels": [C_IN, C_OUT],
                       "parallelism": "dp%d" % world, "l2": "flushed between timed iterations (256 MiB memset)",
                       "launch": ("cuda_graph" if graph is not None else "eager") + ", one stream per (level, branch)",
                       "tflops_per_s": round(flops_step * world / (step_ms * 1e-3) / 1e12, 2),
                       "allreduce_bytes": HEAD_PARAMS * 4 if world > 1 else 0},
            "roofline": roofline,
            "cpu_baseline": {"value": round(cb["value"], 5), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                             "sample": cb["sample"]},
            "e2e": e2e,
            "gpu_launches": int((launches_per_step + torch_fills_per_step) * args.steps),
            "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def measure_e2e(torch, sdb, device, stream, batch, args, world, dist, rank):
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(100 + rank)
    convs = [sdb.DeformConv(C_IN, C_OUT, 3, 1, 1).to(device, bf) for _ in range(2)]
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

    # one stream per FPN level (what a user of the operator API would do for independent maps): the H2D
    # copies of the later levels overlap the kernels of the earlier ones, and the small maps (1-66 tiles)
    # run beside each other.  autograd runs each backward on its forward's stream; weight.grad accumulation
    # across streams is ordered by the autograd engine (leaf created on `stream`).
    side = [torch.cuda.Stream(device) for _ in LEVELS]

    def step():
        main = torch.cuda.current_stream()
        for c in convs:
            c.weight.grad = None
        offs = []
        for lv, st in zip(host, side):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                off = lv["off"].to(device, non_blocking=True).requires_grad_()
                offs.append(off)
                for b in range(2):
                    x = lv["x"][b].to(device, non_blocking=True).requires_grad_()
                    gy = lv["gy"][b].to(device, non_blocking=True)
                    y = convs[b](x, off)
                    y.backward(gy)
        for st in side:
            main.wait_stream(st)
        if world > 1:
            params = {"cls_dcn.weight": convs[0].weight, "refine_dcn.weight": convs[1].weight}
            bucket.pack({k: p.grad for k, p in params.items()})
            bucket.all_reduce(average=True)
            bucket.unpack(params)
        for b in range(2):
            out_host[b].copy_(convs[b].weight.grad, non_blocking=True)
        for i, off in enumerate(offs):
            goff_host[i].copy_(off.grad, non_blocking=True)
        main.synchronize()  # the step's results are on the host

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            step()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
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
            "api": "slenderobjdet_b200.DeformConv.forward + autograd backward, pinned host tensors, one torch stream per FPN level"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cb = cpu_reference(steps=args.steps, warmup=args.warmup, seconds_hint=10.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(cb["value"], 5), "unit": UNIT, "n_gpus": world,
        "steps": cb["steps"], "warmup": args.warmup, "ms_per_step": round(cb["ms_per_step"], 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RepPoints-v1 R-50-FPN head DCN pair, P3-P7 @800x1344, batch %d, fwd+bwd "
                               "(BASELINE.json configs[1]); CPU deform_conv2d path on a bounded sample" % BATCH_PER_GPU,
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
