#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the FlashAttention-forward hot path.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload C2|C1|C3|C4|C5]

One "step" = one forward pass O = softmax(QK^T/sqrt(d)) V over one batch of synthetic N(0,1) inputs.
Default workload = BASELINE.json configs[1] (the README headline shape): B=2 H=8 d=64 N=8192, fp32 in HBM,
tf32 tensor-core contractions.  With N > 1 ranks (torchrun, one process per GPU) the B x H axis is sharded with no
data-path collective and per-GPU work is fixed ("weak"): every rank runs the same 16-head workload.

--workload C5 is the ring-attention config (B=1 H=32 d=128 N=131072 bf16, sequence split over the ranks, strong scaling).

Prints ONE JSON line on rank 0:
  value      TFLOP/s, whole job, kernel timed with CUDA events on the launch stream, inputs resident in HBM,
             L2 flushed (256 MiB write) before every timed step
  e2e        the same metric through the host-buffer C-ABI entry (fa_forward_host): pinned host Q/K/V -> device,
             kernel, O -> host, inside the timed region
  roofline   tensor-pipe roofline of the dominant (only) kernel
  cpu_baseline  torch CPU softmax(QK^T/sqrt(d))V on the box's host cores, bounded sample, rank 0 / N=1 only
  --impl reference: the UNMODIFIED reference CUDA kernel (src/main.cpp + src/flashattention.cu from /root/reference,
             rebuilt for sm_100a into oracle/_ref/flash_ref_d64.so) through its own forward(Q,K,V,causal) on the same
             GPU; if that build is absent the oracle's CPU port is timed instead and labelled so.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {  # name: (B, H, N, d, dtype)
    "C1": (2, 8, 1024, 64, "f32"),
    "C2": (2, 8, 8192, 64, "f32"),
    "C3": (8, 16, 1024, 32, "f32"),
    "C4": (4, 32, 8192, 128, "bf16"),
}
L2_FLUSH_BYTES = 256 << 20


def flops_of(B, H, N, d):
    return 4.0 * B * H * N * N * d


def bytes_of(B, H, N, d, es):
    return 4.0 * B * H * N * d * es


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"bf16": float(j["bf16_tflops"]), "bf16_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "hbm": float(j["hbm_gbs"]), "src": "MEASURED_PEAKS.json"}
    return {"bf16": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_inputs(torch, B, H, N, d, dtype, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    hosts = [torch.randn(B, H, N, d, generator=g, dtype=torch.float32).to(tdt).pin_memory() for _ in range(3)]
    devs = [h.to(device, non_blocking=True) for h in hosts]
    return hosts, devs


def time_kernel(torch, fn, steps, warmup, flush):
    """Per-step CUDA-event timing on the current stream with an L2 flush before every timed step."""
    for _ in range(warmup):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def cpu_baseline(torch, N, d, budget_s=12.0):
    """torch CPU attention (bench_flashattention.py:36-40 with the 1/sqrt(d) scale) on a bounded sample of the workload:
    whole (b,h) slices of the same N and d, as many as fit the time budget (>= 1), chunked over 2048 query rows."""
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(7)
    q, k, v = (torch.randn(N, d, generator=g) for _ in range(3))
    scale = 1.0 / math.sqrt(d)

    def one_head():
        out = torch.empty(N, d)
        for r0 in range(0, N, 2048):
            s = (q[r0:r0 + 2048] @ k.t()) * scale
            out[r0:r0 + 2048] = torch.softmax(s, dim=-1) @ v
        return out

    one_head()  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        one_head()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 16:
            break
    dt = time.perf_counter() - t0
    return {"value": round(flops_of(1, 1, N, d) * n / dt * 1e-12, 4), "unit": "TFLOP/s", "cores": cores, "kind": "port",
            "sample": f"{n} of the workload's (batch, head) slices at full N={N}, d={d}; torch CPU fp32 softmax(QK^T/sqrt(d))V, "
                      f"{dt:.1f} s, scaled linearly"}


def reference_cpu_loop(budget_s=10.0):
    """The reference's own CPU implementation of the path — attention_forward_cpu (src/llm.c/attention_forward.cu:53-125:
    causal, 1/sqrt(hs), scalar single-threaded C) — from oracle/_ref/libllmc_ref.so, timed on a bounded sample.
    None where the reference was not compiled (oracle/_ref absent)."""
    import numpy as np

    from oracle import fa_oracle

    B, T, NH, hs = 1, 1024, 4, 64
    C = NH * hs
    inp = np.random.default_rng(0).random((B, T, 3 * C), dtype=np.float32) * 2 - 1
    if fa_oracle.ref_llmc_cpu(inp, B, T, C, NH) is None:   # warm-up + availability
        return None
    t0 = time.perf_counter()
    n = 0
    while n < 20:
        fa_oracle.ref_llmc_cpu(inp, B, T, C, NH)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / n
    fl = 4.0 * B * NH * hs * T * (T + 1) / 2   # visible (query, key) pairs x 2 contractions x 2 FLOP
    return {"value": round(fl / dt * 1e-12, 6), "unit": "TFLOP/s", "cores": 1, "kind": "reference",
            "sample": f"reference attention_forward_cpu (llm.c CPU loop, causal, scalar, 1 thread): B={B} T={T} NH={NH} hs={hs}, "
                      f"{n} calls of {dt * 1e3:.1f} ms; FLOPs counted over the visible pairs only"}


def run_ours(args, torch, dist, rank, world, device):
    import flashattention_c_b200 as fab

    B, H, N, d, dtype = WORKLOADS[args.workload]
    es = 2 if dtype == "bf16" else 4
    hosts, devs = make_inputs(torch, B, H, N, d, dtype, device, 1234 + rank)
    q, k, v = devs
    scale = 1.0 / math.sqrt(d)
    out = torch.empty_like(q)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
    fn = lambda: fab.attention(q, k, v, causal=False, scale=scale, out=out)  # noqa: E731
    fn()
    torch.cuda.synchronize()
    assert fab.last_impl() == fab.FA_IMPL_TCGEN05, "bench must run the tcgen05 kernel"

    sampler = ClockSampler(torch.cuda.current_device())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = fab.launch_count()
    wall0 = time.perf_counter()
    ms = time_kernel(torch, fn, args.steps, args.warmup, flush)
    launches = fab.launch_count() - launches0 - args.warmup
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    total_ms = sum(ms)

    # e2e: pinned host buffers -> device -> kernel -> host, through the C-ABI host entry
    o_host = torch.empty_like(hosts[0]).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        fab.attention_host(hosts[0], hosts[1], hosts[2], causal=False, scale=scale, out=o_host)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fab.attention_host(hosts[0], hosts[1], hosts[2], causal=False, scale=scale, out=o_host)
    e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = t.tolist()
    if rank != 0:
        return None
    fl = flops_of(B, H, N, d)
    ms_per_step = total_ms / args.steps
    value = fl * world / (ms_per_step * 1e-3) * 1e-12
    peaks = load_peaks()
    peak = peaks["bf16"] if dtype == "bf16" else peaks["bf16"] / 2
    achieved = fl / (ms_per_step * 1e-3) * 1e-12
    line = {
        "metric": "fwd attention TFLOP/s (B2 H8 d64 N8192)" if args.workload == "C2" else f"fwd attention TFLOP/s ({args.workload})",
        "value": round(value, 2), "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if dtype == "f32" else "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: B={B} H={H} d={d} N={N} {'fp32-in/tf32' if dtype == 'f32' else 'bf16'} non-causal, "
                               f"scale=1/sqrt(d), per GPU; B*H sharded across {world} GPU(s), no collective",
                   "l2": "flushed (256 MiB write) before every timed step", "global_bh": B * H * world,
                   "flops_per_step_per_gpu": fl, "algorithmic_bytes_per_step_per_gpu": bytes_of(B, H, N, d, es)},
        "e2e": {"value": round(fl * world / (e2e_s / e2e_steps) * 1e-12, 3), "unit": "TFLOP/s", "ms_per_step": round(e2e_s / e2e_steps * 1e3, 4),
                "h2d_bytes_per_step": 3 * q.numel() * es, "d2h_bytes_per_step": q.numel() * es, "steps": e2e_steps,
                "api": "fa_forward_host (C-ABI, pinned host buffers)"},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s",
                     "frac": round(achieved / peak, 4), "traffic": TRAFFIC_BYTES.get(args.workload),
                     "peak_source": peaks["src"] + (" bf16_tflops / 2 (tf32, derived)" if dtype == "f32" else " bf16_tflops (burst)"),
                     "kernel": "fa_fwd_sm100_kernel", "hbm_gbs_achieved": round(bytes_of(B, H, N, d, es) / (ms_per_step * 1e-3) * 1e-9, 1)},
        "ms_min": round(min(ms), 5), "ms_median": round(sorted(ms)[len(ms) // 2], 5), "wall_s": round(wall, 3),
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(torch, N, d)
        line["cpu_baseline"]["reference_cpu_loop"] = reference_cpu_loop(5.0)   # the reference's own scalar CPU loop, beside it
    if world == 1 and not args.no_extra:
        line["other_configs"] = other_configs(torch, fab, device, flush, peaks)
    return line


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/); None = not captured
TRAFFIC_BYTES = {"C2": 117.4e6, "C4": 1057.8e6}   # dram__bytes_read+write per launch, profiles/r01_ncu_summary.md (v9 kernel)


def other_configs(torch, fab, device, flush, peaks):
    """Kernel-only numbers for the remaining single-GPU BASELINE configs (reported, not the headline)."""
    res = {}
    for name in ("C1", "C3", "C4"):
        B, H, N, d, dtype = WORKLOADS[name]
        es = 2 if dtype == "bf16" else 4
        _, (q, k, v) = make_inputs(torch, B, H, N, d, dtype, device, 99)
        out = torch.empty_like(q)
        scale = 1.0 / math.sqrt(d)
        ms = time_kernel(torch, lambda: fab.attention(q, k, v, scale=scale, out=out), 10, 3, flush)
        med = sorted(ms)[len(ms) // 2]
        peak = peaks["bf16"] if dtype == "bf16" else peaks["bf16"] / 2
        res[name] = {"ms": round(med, 5), "tflops": round(flops_of(B, H, N, d) / med * 1e-9, 1),
                     "frac_tensor_peak": round(flops_of(B, H, N, d) / med * 1e-9 / peak, 4),
                     "hbm_gbs": round(bytes_of(B, H, N, d, es) / med * 1e-6, 1), "frac_hbm_peak": round(bytes_of(B, H, N, d, es) / med * 1e-6 / peaks["hbm"], 4)}
        del q, k, v, out
    return res


def run_ring(args, torch, dist, rank, world, device):
    """--workload C5: B=1 H=32 d=128 N=131072 bf16, the sequence split over the ranks (ring attention; one GPU: the plain
    forward over the whole sequence).  Total work is fixed ("strong" scaling): value = 4*H*N^2*d / max-over-ranks time."""
    import flashattention_c_b200 as fab

    H, N, d = 32, 131072, 128
    n_loc = N // world
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    q, k, v = (torch.randn(1, H, n_loc, d, device=device, generator=g).to(torch.bfloat16) for _ in range(3))
    fn = lambda: fab.ring_attention(q, k, v, causal=False)   # noqa: E731  (p2p transport: copy-engine pulls over NVLink)
    for _ in range(args.warmup):
        o, _lse = fn()
    torch.cuda.synchronize()
    assert fab.last_impl() == fab.FA_IMPL_TCGEN05
    sampler = ClockSampler(torch.cuda.current_device())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = fab.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    launches = fab.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    # e2e: this rank's shards from pinned host memory, the ring forward, this rank's O shard back to pinned host memory
    hq, hk, hv = (t.cpu().pin_memory() for t in (q, k, v))
    ho = torch.empty_like(hq).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dq, dk, dv = (t.to(device, non_blocking=True) for t in (hq, hk, hv))
        o, _lse = fab.ring_attention(dq, dk, dv, causal=False)
        ho.copy_(o, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = t.tolist()
    if rank != 0:
        return None
    fl = 4.0 * H * float(N) * N * d
    ms_per_step = total_ms / args.steps
    peaks = load_peaks()
    achieved = fl / world / (ms_per_step * 1e-3) * 1e-12       # per GPU
    shard_bytes = q.numel() * 2
    return {
        "metric": "fwd attention TFLOP/s (C5: B1 H32 d128 N131072 bf16, ring)", "value": round(fl / (ms_per_step * 1e-3) * 1e-12, 1),
        "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"C5: B=1 H={H} d={d} N={N} bf16 non-causal, scale=1/sqrt(d); sequence split over {world} GPU(s), "
                               f"{n_loc} rows per GPU; K/V shards pulled from their owners by the copy engines (p2p transport) under the kernel",
                   "l2": "inputs larger than L2 (K+V shard per step: %d MB)" % (2 * shard_bytes >> 20),
                   "flops_per_step_total": fl, "kv_bytes_pulled_per_gpu_per_step": 2 * shard_bytes * (world - 1)},
        "e2e": {"value": round(fl / (e2e_s / e2e_steps) * 1e-12, 1), "unit": "TFLOP/s", "ms_per_step": round(e2e_s / e2e_steps * 1e3, 3),
                "h2d_bytes_per_step": 3 * shard_bytes * world, "d2h_bytes_per_step": shard_bytes * world, "steps": e2e_steps,
                "api": "ring_attention on this rank's shards copied from / to pinned host memory"},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        # a step is hundreds of milliseconds of back-to-back tensor work under the power cap: the sustained cuBLAS figure is the
        # denominator (the burst one is given beside it)
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": round(peaks["bf16_sustained"], 1), "unit": "TFLOP/s",
                     "frac": round(achieved / peaks["bf16_sustained"], 4), "traffic": None,
                     "peak_source": peaks["src"] + " bf16_tflops_sustained (per GPU; long step under the power cap)",
                     "frac_of_burst_peak": round(achieved / peaks["bf16"], 4),
                     "kernel": "fa_fwd_sm100_kernel (bf16 d=128, fp32 partial output) x ring steps + fa_merge_kernel"},
    }


def run_reference(args, torch, rank, world, device):
    """The reference arm: its own forward(Q,K,V,causal) (scale fixed at 1.0 inside, src/flashattention.cu:593)."""
    if rank != 0:
        return None
    from oracle import fa_oracle

    long_seq = args.workload == "C5"     # one forward of the reference kernel takes ~15 s at N = 131072: one warm-up, one step
    B, H, N, d, dtype = (1, 32, 131072, 128, "bf16") if long_seq else WORKLOADS[args.workload]
    fl = flops_of(B, H, N, d)
    base = {"metric": "fwd attention TFLOP/s (B2 H8 d64 N8192)" if args.workload == "C2" else f"fwd attention TFLOP/s ({args.workload})",
            "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic", "impl": "reference"}
    ext = fa_oracle.load_ref_torch_ext(d)  # bf16 workloads: the reference runs fp32 copies, it has no bf16 path
    if ext is not None and torch.cuda.is_available():
        g = torch.Generator().manual_seed(1234)
        hosts = [torch.randn(B * H, N, d, generator=g).pin_memory() for _ in range(3)]
        q, k, v = (h.to(device) for h in hosts)
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
        steps = 1 if long_seq else max(1, min(args.steps, 10))
        warm = 1 if long_seq else max(1, min(args.warmup, 3))
        ms = time_kernel(torch, lambda: ext.forward(q, k, v, False), steps, warm, flush)
        ms_per_step = sum(ms) / len(ms)
        o_host = torch.empty(B * H, N, d).pin_memory()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(steps, 3))
        for _ in range(e2e_steps):
            dq, dk, dv = (h.to(device, non_blocking=True) for h in hosts)
            o_host.copy_(ext.forward(dq, dk, dv, False))
            torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        base.update({"value": round(fl / (ms_per_step * 1e-3) * 1e-12, 3), "ms_per_step": round(ms_per_step, 4), "steps": steps, "warmup": warm,
                     "dtype": "f32", "config": {"workload": f"{args.workload}: B={B} H={H} d={d} N={N} fp32 non-causal; reference CUDA kernel "
                                                            "flash_tiled_coarse rebuilt for sm_100a (oracle/_ref), its forward() incl. torch::zeros + cudaDeviceSynchronize, scale fixed at 1.0"},
                     "e2e": {"value": round(fl / e2e_s * 1e-12, 3), "unit": "TFLOP/s", "ms_per_step": round(e2e_s * 1e3, 3),
                             "h2d_bytes_per_step": 3 * q.numel() * 4, "d2h_bytes_per_step": q.numel() * 4},
                     "cpu_baseline": reference_cpu_loop() or {
                         "value": None, "unit": "TFLOP/s", "cores": 0, "kind": "reference",
                         "sample": "the reference for this path is a CUDA kernel; it ran on the GPU, not on host cores"}})
        base["cpu_baseline"]["note"] = ("the reference's implementation of this path is a CUDA kernel: `value` and `e2e` of this line are "
                                        "that kernel on the same GPU; cpu_baseline is the only CPU code the reference has for it")
        return base
    # no reference build on this box: time the oracle's CPU port of the same algorithm on a bounded sample
    import numpy as np

    n_s = min(N, 2048)
    rng = np.random.default_rng(0)
    q, k, v = (rng.standard_normal((1, n_s, d), dtype=np.float32) for _ in range(3))
    fa_oracle.tiled(q, k, v, 1.0, False)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 10.0 and reps < args.steps:
        fa_oracle.tiled(q, k, v, 1.0, False)
        reps += 1
    dt = (time.perf_counter() - t0) / max(reps, 1)
    val = flops_of(1, 1, n_s, d) / dt * 1e-12
    base.update({"value": round(val, 5), "ms_per_step": round(dt * 1e3, 3), "dtype": "f32",
                 "config": {"workload": f"{args.workload} (bounded sample: 1 head, N={n_s}, d={d}); oracle CPU port of the reference recurrence"},
                 "cpu_baseline": {"value": round(val, 5), "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
                                  "sample": f"1 (batch, head) slice, N={n_s}, OpenMP over query tiles"},
                 "e2e": {"value": round(val, 5), "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS) + ["C5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_configs leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps(run_reference(args, torch, 0, world, None)))
            return 0
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    try:
        if args.impl == "ours" and args.workload == "C5":
            line = run_ring(args, torch, dist, rank, world, device)
        elif args.impl == "ours":
            line = run_ours(args, torch, dist, rank, world, device)
        else:
            line = run_reference(args, torch, rank, world, device)
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1 and args.impl == "ours" and dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
