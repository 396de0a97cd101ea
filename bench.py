#!/usr/bin/env python
"""Headline benchmark: co-attention frame-pairs/s on synthetic 60x60x256 feature pairs, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = the co-attention hot path (rgbd_segmentation_RAA.py:150-187 and :204-238, i.e. the RGB and
the depth modality call) over one batch of 32 synthetic frame pairs per GPU.  Prints ONE JSON line
(rank 0).  For N > 1 launch with torchrun (one rank per GPU); pairs are sharded, no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "co-attn frame-pairs/sec @60x60x256"
UNIT = "frame-pairs/s"
C, H, W = 256, 60, 60
L = H * W
PAIRS_PER_GPU = 32
SIGMA = 0.66
# algorithmic flops of one (pair, modality) forward: 6 L^2 C + 2 L C^2; the attend kernel's share
FLOPS_ATTEND_PER_PAIR_MODALITY = 6.0 * L * L * C
FLOPS_PER_PAIR_MODALITY = FLOPS_ATTEND_PER_PAIR_MODALITY + 2.0 * L * C * C


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            pk = json.load(f)
        return {"bf16_tflops": float(pk["bf16_tflops"]), "bf16_tflops_sustained": float(pk.get("bf16_tflops_sustained", 0.0)),
                "hbm_gbs": float(pk["hbm_gbs"]), "source": "measured"}
    # fallback stated in B200_PROFILING.md
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # pragma: no cover - NVML missing
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle side: the only place bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------------
def make_cpu_runner():
    """Returns (run_pair, kind, cores): run_pair() pushes ONE frame pair (RGB + depth modality call, N=1,
    60x60x256, fp32) through the reference's CPU co-attention and returns the seconds spent in the hot path."""
    import torch
    from oracle import coattn_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    v = [torch.from_numpy(x) for x in orc.synthetic_features(1234, 1, H, W, SIGMA, count=4)]
    w_rgb, g_rgb, _ = (None if t is None else torch.from_numpy(t) for t in orc.synthetic_weights(1235, bias=False))
    w_dep, g_dep, b_dep = (torch.from_numpy(t) for t in orc.synthetic_weights(1236, bias=True))
    try:
        from oracle import ref_harness
        if ref_harness.reference_available():
            # the unmodified reference (only present in the build container)
            model = ref_harness.build_stubbed_reference().eval()
            with torch.no_grad():
                model.rgb_similarity_weights.weight.copy_(w_rgb)
                model.gate.weight.copy_(g_rgb.view(1, -1, 1, 1))
                model.depth_similarity_weights.weight.copy_(w_dep)
                model.depth_gate.weight.copy_(g_dep.view(1, -1, 1, 1))
                model.depth_gate.bias.copy_(b_dep)
            marks = {}

            def mark(key):
                def hook(mod, inputs):
                    marks.setdefault(key, time.perf_counter())
                return hook
            model.rgb_similarity_weights.register_forward_pre_hook(mark("rgb0"))
            model.reduce_channels_A.register_forward_pre_hook(mark("rgb1"))
            model.depth_similarity_weights.register_forward_pre_hook(mark("dep0"))
            model.depth_reduce_channels.register_forward_pre_hook(mark("dep1"))

            def run_ref():
                marks.clear()
                with torch.no_grad():
                    ref_harness.run_reference(model, *v)
                # hot-path segments only: W projection .. concat, per modality
                return (marks["rgb1"] - marks["rgb0"]) + (marks["dep1"] - marks["dep0"])
            return run_ref, "reference", cores
    except Exception:
        pass
    from oracle.coattn_torch_cpu import coattention_cpu

    def run_port():
        t0 = time.perf_counter()
        coattention_cpu(v[0], v[1], w_rgb, g_rgb, None)
        coattention_cpu(v[2], v[3], w_dep, g_dep, b_dep)
        return time.perf_counter() - t0
    return run_port, "port", cores


def cpu_baseline(budget_s: float, min_pairs: int = 3, max_pairs: int = 64):
    run_pair, kind, cores = make_cpu_runner()
    run_pair()  # warm-up
    spent, times = 0.0, []
    while (len(times) < min_pairs or spent < budget_s) and len(times) < max_pairs:
        dt = run_pair()
        times.append(dt)
        spent += dt
    med = sorted(times)[len(times) // 2]
    return {"value": 1.0 / med, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(times)} frame pairs (N=1, RGB+depth call each, 60x60x256, fp32, {cores} threads), "
                      f"median {med * 1e3:.1f} ms/pair, best {min(times) * 1e3:.1f} ms/pair"}


def run_reference_arm(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return
    run_pair, kind, cores = make_cpu_runner()
    for _ in range(max(1, args.warmup)):
        run_pair()
    hot = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):   # one step = a bounded sample of the workload: ONE frame pair of the batch
        hot += run_pair()
    elapsed = time.perf_counter() - t0
    value = args.steps / hot
    sample = (f"{args.steps} steps x 1 frame pair (N=1, RGB+depth call, 60x60x256, fp32, {cores} threads); "
              f"hot-path seconds {hot:.2f} of {elapsed:.2f} wall")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": 1e3 * hot / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "co-attention module alone on synthetic 60x60x256 feature pairs (BASELINE cfg 2); "
                               "CPU arm: one frame pair of the batch per step",
                   "feat_hw": [H, W], "channels": C, "sigma": SIGMA},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.nn.functional as F

    from cosnet_b200 import _lib
    from cosnet_b200.coattention import HostPipeline, workspace_bytes
    from cosnet_b200.pair_batcher import shard_range

    world, rank, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    peaks = read_peaks()
    FLAGS = _lib.FLAG_BF16 if args.operands == "bf16" else 0
    if args.softmax16:
        FLAGS |= _lib.FLAG_SOFTMAX16
    if args.kmajor:
        FLAGS |= _lib.FLAG_KMAJOR

    # pair-sharded weak scaling: the global batch is world * 32 pairs, this rank owns a contiguous slice
    total_pairs = PAIRS_PER_GPU * world
    start, n = shard_range(total_pairs, world, rank)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + start)

    def feats():
        x = torch.randn((n, C, H, W), generator=g, device=dev, dtype=torch.float32)
        return F.prelu(x, torch.tensor([0.25], device=dev)) * SIGMA
    v_a, v_b, d_a, d_b = feats(), feats(), feats(), feats()
    k = 1.0 / (C ** 0.5)
    w_rgb = (torch.rand((C, C), generator=g, device=dev) * 2 - 1) * k
    w_dep = (torch.rand((C, C), generator=g, device=dev) * 2 - 1) * k
    g_rgb = torch.randn((C,), generator=g, device=dev) * 0.01
    g_dep = torch.randn((C,), generator=g, device=dev) * 0.01
    b_dep = (torch.rand((1,), generator=g, device=dev) * 2 - 1) * k
    cat = [torch.empty((n, 2 * C, H, W), device=dev) for _ in range(4)]
    lse = torch.empty((2, n, L), device=dev)
    mask = torch.empty((2, n, L), device=dev)
    nbytes = workspace_bytes(n, C, H, W)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream

    attend_events = []

    def modality(va, vb, wt, gw, gb, ca, cb, record):
        gbp = None if gb is None else gb.data_ptr()
        _lib.check(lib.coattn_stage_prep_project(va.data_ptr(), vb.data_ptr(), wt.data_ptr(), wsp, nbytes, n, C, H, W, FLAGS, st),
                   "prep_project")
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        _lib.check(lib.coattn_stage_attend_gate(va.data_ptr(), vb.data_ptr(), ca.data_ptr(), cb.data_ptr(), None,
                                                lse.data_ptr(), mask.data_ptr(),
                                                gw.data_ptr(), gbp, wsp, nbytes, n, C, H, W, FLAGS, st), "attend_gate")
        if record:
            e1.record(stream)
            attend_events.append((e0, e1))

    def step(record=False):
        modality(v_a, v_b, w_rgb, g_rgb, None, cat[0], cat[1], record)
        modality(d_a, d_b, w_dep, g_dep, b_dep, cat[2], cat[3], record)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step(record=True)
        ev1.record(stream)
        torch.cuda.synchronize()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    attend_ms = sum(a.elapsed_time(b) for a, b in attend_events) / max(1, len(attend_events))
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = total_pairs * args.steps / (elapsed_ms * 1e-3)

    # ---------------- the stand-alone gate / sigmoid / concat epilogue (cross-check path; the product fuses it into the
    # attend kernel's drain): HBM-bound, 16 L C bytes per sample and side (read Z and V, write the concat)
    zbuf = torch.empty((2, n, C, L), device=dev)
    _lib.check(lib.coattn_stage_prep_project(v_a.data_ptr(), v_b.data_ptr(), w_rgb.data_ptr(), wsp, nbytes, n, C, H, W, FLAGS, st),
               "prep_project")
    _lib.check(lib.coattn_stage_attend(zbuf.data_ptr(), lse.data_ptr(), wsp, nbytes, n, C, H, W, FLAGS, st), "attend")
    gate_events = []
    for i in range(13):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        _lib.check(lib.coattn_stage_gate(zbuf.data_ptr(), v_a.data_ptr(), v_b.data_ptr(), g_rgb.data_ptr(), None,
                                         cat[0].data_ptr(), cat[1].data_ptr(), n, C, H, W, st), "gate")
        g1.record(stream)
        if i >= 3:
            gate_events.append((g0, g1))
    torch.cuda.synchronize()
    gate_ms = sum(a.elapsed_time(b) for a, b in gate_events) / len(gate_events)
    gate_bytes = 2 * n * 16.0 * L * C
    del zbuf
    # restore cat[0] / cat[1] (the e2e legs compare against them)
    step()
    torch.cuda.synchronize()

    # ---------------- end to end: host buffers in, host buffers out, through the public host API
    e2e_steps = max(2, min(args.steps, 5))
    pipe = HostPipeline(n, C, H, W, chunk=4, slots=3, device=dev, bf16_operands=bool(FLAGS & _lib.FLAG_BF16))
    hin = [t.cpu().pin_memory() for t in (v_a, v_b, d_a, d_b)]
    hout = [torch.empty((n, 2 * C, H, W), dtype=torch.float32).pin_memory() for _ in range(4)]

    def e2e_step():
        pipe(hin[0], hin[1], w_rgb, g_rgb, None, hout[0], hout[1])
        pipe.wait_host()
        pipe(hin[2], hin[3], w_dep, g_dep, b_dep, hout[2], hout[3])
        pipe.wait_host()
    e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_pairs * e2e_steps / e2e_s
    # sanity: the pipelined host path must reproduce the resident path bit for bit
    same = bool(torch.equal(hout[0], cat[0].cpu()) and torch.equal(hout[3], cat[3].cpu()))

    # the same through the gated-only contract (the consumer splits the reduce conv, so the passthrough
    # half of the concat -- a copy of the inputs the host already holds -- is neither produced nor sent back)
    gpipe = HostPipeline(n, C, H, W, chunk=4, slots=3, device=dev, bf16_operands=bool(FLAGS & _lib.FLAG_BF16), gated_only=True)
    gout = [torch.empty((n, C, H, W), dtype=torch.float32).pin_memory() for _ in range(4)]

    def e2e_gated_step():
        gpipe(hin[0], hin[1], w_rgb, g_rgb, None, gout[0], gout[1])
        gpipe(hin[2], hin[3], w_dep, g_dep, b_dep, gout[2], gout[3])
    e2e_gated_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_gated_step()
    torch.cuda.synchronize()
    e2e_gated_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_gated_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_gated_s = float(t.item())
    same_gated = bool(torch.equal(gout[0], cat[0][:, :C].cpu()) and torch.equal(gout[3], cat[3][:, :C].cpu()))

    # ---------------- the 16-bit feature interface (coattn_forward16; SURVEY 8b "fp32 (or bf16)", 8f N4): the same pairs
    # as fp16 features in and out, resident and through the host pipeline with fp16 host buffers.  Reported beside the
    # headline, never instead of it: `value` / `e2e` above keep the reference's fp32 feature contract.
    dt16 = torch.bfloat16 if (FLAGS & _lib.FLAG_BF16) else torch.float16
    F16 = FLAGS & _lib.FLAG_BF16
    f16 = [t.to(dt16) for t in (v_a, v_b, d_a, d_b)]
    c16 = [torch.empty((n, 2 * C, H, W), device=dev, dtype=dt16) for _ in range(4)]

    def step16():
        _lib.check(lib.coattn_forward16(f16[0].data_ptr(), f16[1].data_ptr(), w_rgb.data_ptr(), g_rgb.data_ptr(), None,
                                        c16[0].data_ptr(), c16[1].data_ptr(), None, None, wsp, nbytes, n, 1, C, H, W, F16, st),
                   "coattn_forward16")
        _lib.check(lib.coattn_forward16(f16[2].data_ptr(), f16[3].data_ptr(), w_dep.data_ptr(), g_dep.data_ptr(),
                                        b_dep.data_ptr(), c16[2].data_ptr(), c16[3].data_ptr(), None, None, wsp, nbytes, n, 1,
                                        C, H, W, F16, st), "coattn_forward16")
    io16_steps = max(2, min(args.steps, 50))
    for _ in range(3):
        step16()
    torch.cuda.synchronize()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    for _ in range(io16_steps):
        step16()
    h1.record(stream)
    torch.cuda.synchronize()
    io16_ms = h0.elapsed_time(h1)
    pipe16 = HostPipeline(n, C, H, W, chunk=4, slots=3, device=dev, feature_dtype=dt16)
    hin16 = [t.cpu().pin_memory() for t in f16]
    hout16 = [torch.empty((n, 2 * C, H, W), dtype=dt16).pin_memory() for _ in range(4)]

    def e2e16_step():
        pipe16(hin16[0], hin16[1], w_rgb, g_rgb, None, hout16[0], hout16[1])
        pipe16(hin16[2], hin16[3], w_dep, g_dep, b_dep, hout16[2], hout16[3])
    e2e16_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e16_step()
    torch.cuda.synchronize()
    e2e16_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([io16_ms, e2e16_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        io16_ms, e2e16_s = float(t[0].item()), float(t[1].item())
    same16 = bool(torch.equal(hout16[0], c16[0].cpu()) and torch.equal(hout16[3], c16[3].cpu()))
    # ... and with the gated-only contract on top (16-bit, no passthrough half on the wire): the least PCIe traffic per pair
    gpipe16 = HostPipeline(n, C, H, W, chunk=4, slots=3, device=dev, feature_dtype=dt16, gated_only=True)
    gout16 = [torch.empty((n, C, H, W), dtype=dt16).pin_memory() for _ in range(4)]

    def e2e16_gated_step():
        gpipe16(hin16[0], hin16[1], w_rgb, g_rgb, None, gout16[0], gout16[1])
        gpipe16(hin16[2], hin16[3], w_dep, g_dep, b_dep, gout16[2], gout16[3])
    e2e16_gated_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e16_gated_step()
    torch.cuda.synchronize()
    e2e16_gated_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e16_gated_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e16_gated_s = float(t.item())
    same16_gated = bool(torch.equal(gout16[0], c16[0][:, :C].cpu()) and torch.equal(gout16[3], c16[3][:, :C].cpu()))
    # the 16-bit results are the fp32-interface results rounded once (fp32 features that are not 16-bit values differ in the
    # passthrough half only by that rounding): report the distance instead of asserting bit equality here
    rel16 = float(((c16[0].float() - cat[0]).norm() / cat[0].norm()).item())

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    achieved_tflops = n * FLOPS_ATTEND_PER_PAIR_MODALITY / (attend_ms * 1e-3) / 1e12
    # Denominator: when the attend kernel is timed INSIDE a long step loop the GPU sits at its power cap (SM clock
    # ~1.5 GHz, reason sw_power_cap) and the sustained bf16 peak of MEASURED_PEAKS.json applies; a short run that never
    # left the boost clock is held against the burst peak.  The fraction of the burst peak is always kept beside it.
    clk = clocks.summary()
    capped = bool(clk.get("sm_mhz") and clk.get("sm_max_mhz") and clk["sm_mhz"] < 0.9 * clk["sm_max_mhz"])
    sustained = (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]) if capped else peaks["bf16_tflops"]
    roofline = {
        "kernel": "attend2_kernel", "bound": "tensor", "achieved": achieved_tflops, "peak": sustained,
        "unit": "TFLOP/s", "frac": achieved_tflops / sustained,
        # dram__bytes_read.sum + dram__bytes_write.sum of one attend2 launch at this shape, from the committed
        # `ncu --set full` capture (profiles/r1_ncu_kernels.txt): 440.3 MB read + 427.9 MB written
        "traffic": 868.3e6 if (n == PAIRS_PER_GPU and (FLAGS & _lib.FLAG_BF16) == 0) else None,
        "peak_kind": (f"{peaks['source']} sustained dense bf16 (kernel timed inside the {args.steps}-step loop, SM clock "
                      f"{clk.get('sm_mhz')} MHz under the power cap); burst peak {peaks['bf16_tflops']}") if capped else
                     f"{peaks['source']} burst dense bf16 (clocks stayed at boost)",
        "frac_of_burst_peak": achieved_tflops / peaks["bf16_tflops"],
        "ms_per_launch": attend_ms, "algorithmic_flops_per_launch": n * FLOPS_ATTEND_PER_PAIR_MODALITY,
        "executed_flops_per_launch": n * FLOPS_ATTEND_PER_PAIR_MODALITY * 8.0 / 6.0,
        "whole_step_frac": (2 * n * FLOPS_PER_PAIR_MODALITY / (elapsed_ms / args.steps * 1e-3) / 1e12) / sustained,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.operands + " operands, f32 accumulate", "data": "synthetic",
        "config": {
            "workload": "co-attention module alone on synthetic 60x60x256 feature pairs, batch 32 per GPU "
                        "(BASELINE cfg 2); one step = RGB + depth modality call over the batch",
            "pairs_per_gpu": PAIRS_PER_GPU, "feat_hw": [H, W], "channels": C, "sigma": SIGMA,
            "parallelism": f"pair-sharded x{world}, no data-path collective",
            "l2": f"inputs of one step ({4 * n * C * L * 4 / 1e6:.0f} MB) exceed the 126 MB L2; no explicit flush",
            "flops_per_frame_pair": 2 * FLOPS_PER_PAIR_MODALITY,
        },
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * pipe.h2d_bytes * world,
                "d2h_bytes_per_step": 2 * pipe.d2h_bytes * world, "steps": e2e_steps, "matches_resident_path": same,
                "api": "cosnet_b200.coattention.HostPipeline (pinned host in/out, 3 streams x chunks of 4 pairs)",
                "gated_only_contract": {"value": total_pairs * e2e_steps / e2e_gated_s, "unit": UNIT,
                                        "d2h_bytes_per_step": 2 * gpipe.d2h_bytes * world, "matches_resident_path": same_gated,
                                        "note": "outputs [n,256,h,w]: the concat's passthrough half (= the caller's own inputs) "
                                                "is not sent back; for the split-reduce-conv consumer"}},
        "io16": {"value": total_pairs * io16_steps / (io16_ms * 1e-3), "unit": UNIT, "ms_per_step": io16_ms / io16_steps,
                 "steps": io16_steps, "feature_dtype": str(dt16), "gpu_launches_per_step": 6,
                 # whole modality call (cast_w + project_mn + attend2), algorithmic flops 6 L^2 C + 2 L C^2 per pair
                 "whole_call_tflops": 2 * n * FLOPS_PER_PAIR_MODALITY / (io16_ms / io16_steps * 1e-3) / 1e12,
                 "whole_call_frac_of_burst_peak": 2 * n * FLOPS_PER_PAIR_MODALITY / (io16_ms / io16_steps * 1e-3) / 1e12 / peaks["bf16_tflops"],
                 "rel_l2_vs_fp32_interface": rel16,
                 "e2e": {"value": total_pairs * e2e_steps / e2e16_s, "unit": UNIT,
                         "h2d_bytes_per_step": 2 * pipe16.h2d_bytes * world, "d2h_bytes_per_step": 2 * pipe16.d2h_bytes * world,
                         "matches_resident_path": same16,
                         "gated_only_contract": {"value": total_pairs * e2e_steps / e2e16_gated_s, "unit": UNIT,
                                                 "d2h_bytes_per_step": 2 * gpipe16.d2h_bytes * world,
                                                 "matches_resident_path": same16_gated}},
                 "note": "coattn_forward16: 16-bit features in and out (host buffers 16-bit as well), operands read in place "
                         "by TMA, no cast pass; outside the headline's timed region"},
        "gpu_launches": 8 * args.steps,   # per modality call: cast(V_a, V_b), cast_w, project_mn, attend2(+gate+concat)
        "roofline": roofline,
        "epilogue_roofline": {
            "kernel": "gate_kernel (stand-alone gate/sigmoid/scale/concat; the default path fuses it into attend2's drain)",
            "bound": "hbm", "achieved": gate_bytes / (gate_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gate_bytes / (gate_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms_per_launch": gate_ms,
            "algorithmic_bytes_per_launch": gate_bytes, "launches_outside_timed_region": 13,
        },
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(args.cpu_budget)
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--operands", default="f16", choices=["f16", "bf16"], help="16-bit tensor-core operand format")
    ap.add_argument("--kmajor", action="store_true", help="position-major (transposed) operand copies instead of the default "
                    "channel-major planes (cross-check path)")
    ap.add_argument("--softmax16", action="store_true", help="attend kernel with 16 instead of 8 softmax warps (cross-check)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
