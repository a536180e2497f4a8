#!/usr/bin/env python
"""bench.py -- MSDeformAttn forward+backward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload a2d|ytvos|decoder|encoder-a2d|encoder-ytvos] [--regime init|uniform] [--dtype f32|bf16]

A *step* is one pass of the hot path -- MSDeformAttnFunction forward + backward (grad_value,
grad_sampling_loc, grad_attn_weight) -- over one batch of synthetic input.  At N=1 the batch is
BASELINE.json configs[1]: the A2D ResNet-101 encoder shape (5 frames, 360x640 -> levels 45x80, 23x40,
12x20, 6x10; S = Lq = 4820; M=8, D=32, L=4, P=4), fp32.  With N GPUs every rank runs that batch on its
own shard of frames (weak scaling, no collective on the op -- SURVEY.md section 8e).

One JSON line is printed by rank 0:
  value         queries/s (a query = one (n, q) pair, all heads), whole job, inputs resident in HBM,
                device-timed with CUDA events over exactly K steps, max over ranks.  Steps cycle through
                `input_sets` distinct input sets (> L2 in total) so no step finds its inputs in L2.
  e2e           the same metric through the public API with HOST buffers: pinned-host -> device copies of
                value / sampling_locations / attention_weights / grad_output, forward, backward, and
                device -> host copies of output and the three gradients, every step, all inside the timed
                region (steps pipelined over three streams; wall clock around a synchronised region).
  roofline      for the dominant kernel (the backward): algorithmic bytes per launch / its mean launch
                duration (CUDA events around back-to-back graph replays of that kernel alone) vs the measured HBM
                peak (MEASURED_PEAKS.json).
  cpu_baseline  the reference's CPU path (grid_sample formulation, oracle/grid_sample_port.py) timed on
                this box's host cores, rank 0, N=1 only.
`--impl reference` times only that CPU path (all host threads), same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "msdeformattn_fwd_bwd_queries_per_sec"
UNIT = "queries/s"
HBM_FALLBACK_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="a2d", choices=["a2d", "ytvos", "decoder", "encoder-a2d", "encoder-ytvos"],
                    help="a2d (default) = BASELINE.json configs[1]; encoder-* = the 6-layer encoder fwd+bwd with the NCCL "
                         "all-reduce of weight gradients (configs[2], [4]; tools/bench_encoder.py, its own metric)")
    ap.add_argument("--frames-per-gpu", type=int, default=0, help="encoder-* only")
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tf32", "bf16"], help="encoder-* only: GEMM precision policy")
    ap.add_argument("--regime", default="init", choices=["init", "uniform", "module_init"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--input-sets", type=int, default=6)
    ap.add_argument("--no-graph", action="store_true", help="launch from Python instead of CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-CUDA-op leg (oracle/_ref)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget for the cpu_baseline leg")
    return ap.parse_args()


def pick_workload(name):
    from ocpg_b200 import workloads as W
    return {"a2d": W.A2D_ENCODER, "ytvos": W.YTVOS_ENCODER, "decoder": W.A2D_DECODER}[name]


def workload_config(wl, regime):
    """The `config` object of the JSON line: the workload only, identical for both arms (run details go to `run`)."""
    return {"workload": wl.name, "regime": regime, "frames_per_gpu": wl.n_frames, "levels": [list(l) for l in wl.levels],
            "Lq": wl.n_queries, "M": wl.n_heads, "D": wl.head_dim, "L": wl.L, "P": wl.n_points}


def measured_traffic(wl_name, regime, dtype):
    """DRAM bytes per launch of the two kernels from the committed ncu capture (profiles/traffic.json), or None.
    An entry is used only if it was captured from the kernel sources that are built now (`kernel_sources` fingerprint,
    ocpg_b200.source_fingerprint()): a stale capture yields {"stale": ...} and `roofline.traffic` null.
    ``measured_traffic("_on_chip", None, None)`` returns the measured on-chip ceilings stored next to them."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f)
        if wl_name.startswith("_"):
            return table.get(wl_name)
        entry = table.get(f"{wl_name}|{regime}|{dtype}")
        if entry is None:
            return None
        import ocpg_b200
        now = ocpg_b200.source_fingerprint()
        if entry.get("kernel_sources") != now:
            return {"stale": f"ncu capture is of kernel sources {entry.get('kernel_sources')}, built now: {now}"}
        return entry
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (every ~2 ms; the timed
    region of the default run is ~70 ms, too short for `nvidia-smi -lms`), `nvidia-smi` only as a fallback."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        import threading
        self.sm, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:                       # no NVML: one nvidia-smi query at stop()
            self.err = repr(e)
            self.index = index

    def _sample(self):
        self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        for name, bit in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.002)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            try:
                self._sample()
            except Exception:
                pass
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml thread, 2 ms"}
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "samples": 1,
                    "reasons": sorted(n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")),
                    "source": "nvidia-smi after the timed region (NVML unavailable: %s)" % self.err}
        except Exception as e:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable: %r" % e]}


# ----------------------------------------------------------------------------------------------
# CPU reference path (oracle/ -- the one place bench.py may execute it)
# ----------------------------------------------------------------------------------------------
def time_cpu_reference(wl, regime, steps, warmup, budget_s):
    """grid_sample formulation of the reference (ms_deform_attn_func.py:41-61), forward + autograd
    backward, fp32, all host threads.  Each step is a bounded sample of the workload: as many of its
    frames as fit the time budget."""
    import dataclasses
    import torch
    from ocpg_b200.workloads import make_inputs
    from oracle.grid_sample_port import msda_grid_sample_fwd_bwd

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = make_inputs(wl, regime, seed=0, device="cpu")

    def run(nf):
        msda_grid_sample_fwd_bwd(x["value"][:nf], x["shapes"], x["loc"][:nf], x["attn"][:nf], x["grad_out"][:nf])

    t0 = time.perf_counter(); run(1); t1 = time.perf_counter() - t0        # first call: page-in, 1 frame
    t0 = time.perf_counter(); run(1); t1 = min(t1, time.perf_counter() - t0)
    per_frame = t1
    nf = wl.n_frames
    while nf > 1 and per_frame * nf * (steps + warmup) > budget_s:
        nf -= 1
    for _ in range(warmup):
        run(nf)
    t0 = time.perf_counter()
    for _ in range(steps):
        run(nf)
    dt = time.perf_counter() - t0
    q = nf * wl.n_queries * steps
    return {"value": q / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nf} of {wl.n_frames} frames per step x {steps} steps (+{warmup} warm-up), fwd+bwd fp32, "
                      f"grid_sample formulation, {torch.get_num_threads()} threads",
            "ms_per_step": dt / steps * 1e3, "frames_per_step": nf}


def time_reference_cuda_op(sets, R, iters, wl):
    """Forward and backward of the reference's CUDA op rebuilt for sm_100a (oracle/_ref/MultiScaleDeformableAttention_ref.so),
    separately, on the same rotating input sets; eager launches queued behind a GPU spin so that no host latency sits
    between the events.  Returns None-like dict with `unavailable` when the .so is not there."""
    import torch
    try:
        from oracle import build_ref_cuda
        if not os.path.exists(build_ref_cuda.SO):
            return {"unavailable": "oracle/_ref/MultiScaleDeformableAttention_ref.so not built (needs /root/reference at build time)"}
        ref = build_ref_cuda.load()
    except Exception as e:
        return {"unavailable": repr(e)[:200]}

    def timed(fn):
        for i in range(3):
            fn(sets[i % R])
        torch.cuda.synchronize()
        evs = []
        torch.cuda._sleep(int(3e7))
        for i in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(sets[i % R]); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in evs]
        return statistics.mean(ts), min(ts)

    f_ms, f_min = timed(lambda x: ref.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64))
    b_ms, b_min = timed(lambda x: ref.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64))
    return {"kind": "reference CUDA op (ms_deform_im2col_cuda.cuh:237-403) rebuilt unmodified for sm_100a, incl. its at::zeros",
            "fwd_ms": f_ms, "fwd_ms_min": f_min, "bwd_ms": b_ms, "bwd_ms_min": b_min,
            "queries_per_s": wl.queries / ((f_ms + b_ms) * 1e-3), "unit": UNIT, "iters": iters,
            "timing": "CUDA events around each eager call, launches queued behind a GPU spin, same rotating input sets"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    wl = pick_workload(args.workload)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes whatever K and W are
    r = time_cpu_reference(wl, args.regime, steps, warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, args.regime),
        "run": {"frames_per_step": r["frames_per_step"], "note": "each step is a bounded sample of the workload's frames (cpu_baseline.sample)"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import ocpg_b200
    from ocpg_b200 import dist as D
    from ocpg_b200 import MultiScaleDeformableAttention as MSDA
    from ocpg_b200.workloads import make_inputs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    rank, local_rank, world = D.init("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    local_cpus = D.bind_to_gpu_cpus(local_rank) if world > 1 else None    # NUMA-local pinned buffers for the e2e leg
    ocpg_b200.lib()                                                   # fail loudly if the extension is missing
    wl = pick_workload(args.workload)
    vdt = torch.bfloat16 if args.dtype == "bf16" else None
    vbytes = 2 if args.dtype == "bf16" else 4
    K, Wm = max(1, args.steps), max(3, args.warmup)
    R = max(1, args.input_sets)

    # R distinct input sets; each rank owns its own shard of frames (seeded by rank): weak scaling
    sets = [make_inputs(wl, args.regime, seed=1000 * rank + i, device=dev, value_dtype=vdt) for i in range(R)]
    fwd_bytes, bwd_bytes = wl.algorithmic_bytes(vbytes, vbytes)
    set_bytes = sum(t.numel() * t.element_size() for t in sets[0].values())

    def step(x):
        out = MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
        grads = MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
        return out, grads

    # one CUDA graph per input set (forward kernel, grad_value memset, backward kernel)
    graphs = None
    if not args.no_graph:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for x in sets:
                step(x)
        torch.cuda.synchronize()
        graphs, keep = [], []
        for x in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep.append(step(x))
            graphs.append(g)

    def run_step(i):
        if graphs is not None:
            graphs[i % R].replay()
        else:
            step(sets[i % R])

    for i in range(Wm):
        run_step(i)
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, barrier + sync on both sides, device-timed, max over ranks
    n0 = ocpg_b200.launch_count()
    clocks = ClockSampler(local_rank)
    D.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        run_step(i)
    e1.record()
    torch.cuda.synchronize(); D.barrier()
    ms_total = D.max_over_ranks(e0.elapsed_time(e1), dev)
    clk = clocks.stop()
    launches = (2 * K) if graphs is not None else (ocpg_b200.launch_count() - n0)
    ms_per_step = ms_total / K
    value_qps = wl.queries * world / (ms_per_step * 1e-3)

    # ---- per-kernel durations: the same rotating inputs, one CUDA graph per (kernel, input set) so that no host
    # launch latency sits between the events and the kernel; mean over `iters` back-to-back replays between one
    # event pair, min over single replays.
    def time_kernel(fn, iters):
        if args.no_graph:
            evs = []
            torch.cuda._sleep(int(3e7))    # ~15 ms of GPU spin: the launches below are queued before the GPU gets to them
            for i in range(iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(sets[i % R]); b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            ts = [a.elapsed_time(b) for a, b in evs]
            return statistics.mean(ts), min(ts)
        gs, keep_ = [], []
        with torch.cuda.stream(torch.cuda.Stream()):
            fn(sets[0])
        torch.cuda.synchronize()
        for x in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep_.append(fn(x))
            gs.append(g)
        for i in range(R):
            gs[i].replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            gs[i % R].replay()
        b.record()
        torch.cuda.synchronize()
        mean = a.elapsed_time(b) / iters
        singles = []
        for i in range(min(iters, 12)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gs[i % R].replay(); b.record()
            torch.cuda.synchronize()
            singles.append(a.elapsed_time(b))
        return mean, min(singles)

    iters = min(K, 60)
    fwd_ms, fwd_min = time_kernel(lambda x: MSDA.ms_deform_attn_forward(
        x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64), iters)
    bwd_ms, bwd_min = time_kernel(lambda x: MSDA.ms_deform_attn_backward(
        x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64), iters)
    # warm-L2 row ("report both, grade on cold"): the same input set replayed back to back, so a
    # launch finds whatever fits of its inputs (A2D: all 86 MB of the forward's) in the 126 MB L2
    warm = None
    if graphs is not None:
        def time_warm(fn):
            with torch.cuda.stream(torch.cuda.Stream()):
                fn(sets[0])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep_w = fn(sets[0])
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            del keep_w
            return a.elapsed_time(b) / iters
        wf = time_warm(lambda x: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64))
        wb = time_warm(lambda x: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64))
        warm = {"fwd_ms": wf, "bwd_ms": wb, "queries_per_s": wl.queries / ((wf + wb) * 1e-3),
                "note": "one input set replayed back to back (inputs partly resident in L2); not the graded row"}

    # ---- the kernel to beat (SURVEY.md section 8d "second GPU baseline"): the reference's own CUDA op
    # (ms_deform_im2col_cuda.cuh:237-403) rebuilt unmodified for sm_100a into oracle/_ref/ (oracle/build_ref_cuda.py), same
    # rotating inputs, device-timed; outside the timed region of `value`, rank 0, fp32 only.  Test infrastructure used as
    # a labelled baseline, never on the product path.
    gpu_baseline = None
    if rank == 0 and args.dtype == "f32" and not args.no_gpu_baseline:
        gpu_baseline = time_reference_cuda_op(sets, R, iters, wl)

    peak, peak_src = hbm_peak()
    traffic = measured_traffic(wl.name, args.regime, args.dtype) or {}
    traffic_note = traffic.pop("stale", None) if isinstance(traffic, dict) else None
    bwd_gbs = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = fwd_bytes / (fwd_ms * 1e-3) / 1e9
    step_gbs = (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers in, host buffers out, through the autograd operator (the call a user makes).
    # Every step copies ITS inputs from pinned host memory and ITS four results back to pinned host memory inside
    # the timed region.  Steps are software-pipelined over three streams (H2D of step i+1 | compute of step i |
    # D2H of step i-1) with two device buffer sets: PCIe is full duplex, so the step costs max(copy in, copy out,
    # compute) instead of their sum.
    e2e = None
    if not args.no_e2e:
        from ocpg_b200 import MSDeformAttnFunction
        hx = {k: sets[0][k].cpu().pin_memory() for k in ("value", "loc", "attn", "grad_out")}
        h2d = sum(t.numel() * t.element_size() for t in hx.values())
        x0 = sets[0]
        outs_shape = [(x0["grad_out"].shape, x0["grad_out"].dtype), (x0["value"].shape, x0["value"].dtype),
                      (x0["loc"].shape, x0["loc"].dtype), (x0["attn"].shape, x0["attn"].dtype)]
        NB = 2
        hout = [[torch.empty(s, dtype=dt).pin_memory() for s, dt in outs_shape] for _ in range(NB)]
        d2h = sum(t.numel() * t.element_size() for t in hout[0])
        dbuf = [{k: torch.empty_like(sets[0][k]) for k in hx} for _ in range(NB)]
        s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        ev_in = [torch.cuda.Event() for _ in range(NB)]
        ev_cmp = [torch.cuda.Event() for _ in range(NB)]
        ev_out = [torch.cuda.Event() for _ in range(NB)]
        keep = [None] * NB

        def e2e_step(i):
            b = i % NB
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_cmp[b])                    # buffer b's previous compute has consumed its inputs
                for k in hx:
                    dbuf[b][k].copy_(hx[k], non_blocking=True)
                ev_in[b].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[b])
                s_cmp.wait_event(ev_out[b])                   # buffer b's previous results have left the device
                v = dbuf[b]["value"].detach().requires_grad_(True)
                sl = dbuf[b]["loc"].detach().requires_grad_(True)
                aw = dbuf[b]["attn"].detach().requires_grad_(True)
                out = MSDeformAttnFunction.apply(v, x0["shapes"], x0["start"], sl, aw, 64)
                out.backward(dbuf[b]["grad_out"])
                keep[b] = (out.detach(), v.grad, sl.grad, aw.grad)
                ev_cmp[b].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[b])
                for h, dv in zip(hout[b], keep[b]):
                    h.copy_(dv, non_blocking=True)
                ev_out[b].record(s_out)

        Ke = min(K, 40)
        torch.cuda.synchronize()
        for i in range(4):
            e2e_step(i)
        D.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(Ke):
            e2e_step(i)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3            # host clock around a fully synchronised region
        D.barrier()
        e2e_ms = D.max_over_ranks(wall_ms, dev) / Ke
        ref_out = MSDA.ms_deform_attn_forward(x0["value"], x0["shapes"], x0["start"], x0["loc"], x0["attn"], 64)
        assert torch.equal(hout[(Ke - 1) % NB][0], ref_out.cpu()), "e2e pipeline returned a different output"
        e2e = {"value": wl.queries * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": Ke,
               "pipeline": "3 streams (H2D | fwd+bwd | D2H), 2 buffer sets, pinned host memory",
               "cpu_affinity": (f"{len(local_cpus)} cores local to the GPU (NVML)" if local_cpus else "unchanged")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = time_cpu_reference(wl, args.regime, steps=3, warmup=1, budget_s=args.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    bwd_kernel = "msda_bwd_sorted" if ocpg_b200.lib().msda_kernel_plan(vbytes, wl.n_heads, wl.head_dim, wl.L, wl.n_points) == 1 \
        and wl.L <= 4 and wl.n_queries == wl.S and wl.n_frames * wl.n_heads * ((wl.n_queries + 31) // 32) > 2 * torch.cuda.get_device_properties(dev).multi_processor_count \
        else "msda_bwd_tiled"
    if rank == 0:
        # The two on-chip resources that actually bound the kernels, from counters of the committed ncu
        # capture and this run's launch times: informative, next to the contractual HBM roofline above.
        on_chip = None
        if traffic.get("bwd_red_bytes") and traffic.get("fwd_l1_lines"):
            oc = measured_traffic("_on_chip", None, None) or {}
            sms, mhz = torch.cuda.get_device_properties(dev).multi_processor_count, (clk or {}).get("sm_mhz") or 1965.0
            red_gbs = traffic["bwd_red_bytes"] / (bwd_ms * 1e-3) / 1e9
            line_rate = traffic["fwd_l1_lines"] / (fwd_ms * 1e-3) / (sms * mhz * 1e6)
            on_chip = {"bwd": {"bound": "SM load/store data path (shared memory + L1: 128 B per clock per SM) -- the row-major kernel's limit; "
                                        "its reds use the L2 fp32 atomic units at `l2_atomic_frac`",
                               "lsu_data_pipe_busy_pct": traffic.get("bwd_lsu_data_pipe_pct"), "red_sector_gbs": red_gbs,
                               "l2_atomic_peak_gbs": oc.get("l2_fp32_atomic_peak_gbs"),
                               "l2_atomic_frac": red_gbs / oc["l2_fp32_atomic_peak_gbs"] if oc.get("l2_fp32_atomic_peak_gbs") else None},
                       "fwd": {"bound": "SM load/store data path: L1 line rate of the gather (without the per-point broadcasts)", "achieved": line_rate,
                               "peak": oc.get("l1_lines_per_clk_per_sm"), "unit": "128-byte lines per clock per SM", "frac": line_rate,
                               "lsu_data_pipe_busy_pct": traffic.get("fwd_lsu_data_pipe_pct")},
                       "source": (oc.get("source") or "") + "; lsu_data_pipe_busy_pct = l1tex__data_pipe_lsu_wavefronts of the committed ncu capture"}
        line = {
            "metric": METRIC, "value": value_qps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(wl, args.regime),
            "run": {"parallelism": f"frames sharded over {world} GPU(s), no collective",
                    "l2_policy": f"steps cycle through {R} distinct input sets of {set_bytes / 1e6:.0f} MB "
                                 f"(+ as much output) each: inputs larger than L2 between reuses",
                    "launch": "python" if graphs is None else "cuda-graph per step (fwd kernel, memset, bwd kernel)"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": bwd_kernel + " (+ the zero-fill of grad_value it needs)", "achieved": bwd_gbs, "peak": peak, "unit": "GB/s",
                         "frac": bwd_gbs / peak, "traffic": traffic.get("bwd"), "traffic_source": traffic.get("source") or traffic_note,
                         "peak_source": peak_src,
                         "algorithmic_bytes": bwd_bytes, "launch_ms": bwd_ms, "launch_ms_min": bwd_min},
            "roofline_fwd": {"kernel": "msda_fwd_tiled", "achieved": fwd_gbs, "frac": fwd_gbs / peak, "traffic": traffic.get("fwd"),
                             "algorithmic_bytes": fwd_bytes, "launch_ms": fwd_ms, "launch_ms_min": fwd_min},
            "roofline_step": {"achieved": step_gbs, "frac": step_gbs / peak, "algorithmic_bytes": fwd_bytes + bwd_bytes},
            "on_chip_ceilings": on_chip, "warm_l2": warm, "gpu_baseline": gpu_baseline,
            "cpu_baseline": cpu, "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    D.finalize()
    return 0


def main():
    args = parse_args()
    if args.workload.startswith("encoder-"):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_encoder
        argv = ["--shape", args.workload.split("-", 1)[1], "--steps", str(args.steps if args.steps != 200 else 10),
                "--warmup", str(min(args.warmup, 5)), "--gemm", args.gemm]
        if args.frames_per_gpu:
            argv += ["--frames-per-gpu", str(args.frames_per_gpu), "--micro", "16"]
        return bench_encoder.main(argv)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
