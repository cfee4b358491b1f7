#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the FlashAttention-forward hot path.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload C2|C1|C3|C4|C5]

One "step" = one forward pass O = softmax(QK^T/sqrt(d)) V over one batch of synthetic N(0,1) inputs.
Default workload = BASELINE.json configs[1] (the README headline shape): B=2 H=8 d=64 N=8192, fp32 in HBM,
tf32 tensor-core contractions.  With N > 1 ranks (torchrun, one process per GPU) the B x H axis is sharded with no
data-path collective and per-GPU work is fixed ("weak"): every rank runs the same 16-head workload.

--workload C5 is the ring-attention config (B=1 H=32 d=128 N=131072 bf16, sequence split over the ranks, strong scaling).

With N > 1 the same JSON line also carries the multi-GPU workloads BASELINE.json names, each with a `parity` field:
  c4_sharded / c3_sharded   configs 4 / 3 at their GLOBAL size (B*H = 128 heads) split over the N ranks, no collective (strong
             scaling): ms (max over ranks), total TFLOP/s, the same problem on one GPU in the same run, efficiency t1/(N*tN);
             parity = every rank's shard bit-equal to the unsharded forward (FA_FLAG_BATCH_INVARIANT) + sampled rows vs fp64
  c5_ring    config 5 (N=131072 split over the ranks): ms, total and per-GPU TFLOP/s, fraction of the sustained bf16 peak,
             overlap (same kernels without transfers / ring), the NCCL transport beside the p2p one; parity = the ring's
             shard vs ONE single-GPU kernel over the full sequence + sampled rows vs fp64

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
HOST_BINDING = {"policy": "unset"}     # filled by main(): how this rank was bound to cores / a NUMA node


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


def pin_to_gpu_numa(local_rank, world):
    """Bind this process (and so the pinned host buffers it allocates afterwards: first-touch) to a private slice of the
    cores of a NUMA node.  FA_BENCH_NUMA=local (default): the node the GPU hangs off (sysfs numa_node of its PCI function);
    spread: nodes round-robin over the local ranks, so that N ranks' host copies do not all pull from one node's DRAM;
    off: leave the affinity alone.  Returns a description for the JSON line."""
    policy = os.environ.get("FA_BENCH_NUMA", "local")
    info = {"policy": policy}
    if policy == "off" or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        import pynvml

        pynvml.nvmlInit()
        nodes = {}
        for d_ in sorted(Path("/sys/devices/system/node").glob("node[0-9]*")):
            cpus = []
            for part in (d_ / "cpulist").read_text().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus += list(range(int(a), int(b or a) + 1))
            if cpus:
                nodes[int(d_.name[4:])] = cpus
        allowed = set(os.sched_getaffinity(0))
        nodes = {n: [c for c in cs if c in allowed] for n, cs in nodes.items()}
        nodes = {n: cs for n, cs in nodes.items() if cs}
        if not nodes:
            return info

        def gpu_node(i):
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            f = Path("/sys/bus/pci/devices") / bus.lower()[-12:] / "numa_node"
            n = int(f.read_text()) if f.exists() else -1
            return n if n in nodes else sorted(nodes)[0]

        order = sorted(nodes)
        node_of = [order[i % len(order)] if policy == "spread" else gpu_node(i) for i in range(world)]
        mine = node_of[local_rank]
        peers = [i for i in range(world) if node_of[i] == mine]
        cpus = nodes[mine]
        per = max(1, len(cpus) // len(peers))
        k = peers.index(local_rank)
        sl = cpus[k * per:(k + 1) * per] or cpus
        os.sched_setaffinity(0, sl)
        info.update({"numa_node": mine, "gpu_numa_node": gpu_node(local_rank), "cpus": f"{sl[0]}-{sl[-1]}", "n_cpus": len(sl),
                     "numa_nodes": len(nodes)})
    except Exception as e:  # affinity is an optimisation: never fail the bench over it
        info["error"] = repr(e)[:120]
    return info


def copy_floor_ms(torch, hosts, o_host, devs, reps=5):
    """Raw pinned-host <-> device copies of exactly the e2e step's bytes (Q, K, V in on one stream, O out on another, in
    parallel): the host/PCIe ceiling under which no e2e time can go.  Caller brackets it with barriers so that all ranks copy
    at the same time."""
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    o_dev = torch.empty_like(devs[0])
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            for h, d_ in zip(hosts, devs):
                d_.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_out):
            o_host.copy_(o_dev, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best


def fp64_rows_check(torch, q, k, v, o, lse, scale, rows, seed):
    """Checker (torch fp64 on the GPU, not the product): `rows` sampled query rows of every (batch, head) of q [B,H,n_q,d]
    against ALL keys of k, v [B,H,n_k,d]; returns (max |O - ref|, max |LSE - ref|)."""
    B, H, n_q, d = q.shape
    g = torch.Generator(device="cpu").manual_seed(seed)
    idx = torch.randperm(n_q, generator=g)[:rows].sort().values.to(q.device)
    err_o = err_l = 0.0
    for h0 in range(0, H, 4):
        qs = q[:, h0:h0 + 4, idx].double()
        s = (qs @ k[:, h0:h0 + 4].double().transpose(-1, -2)) * scale
        ref_l = torch.logsumexp(s, dim=-1)
        ref_o = torch.softmax(s, dim=-1) @ v[:, h0:h0 + 4].double()
        err_o = max(err_o, float((o[:, h0:h0 + 4, idx].double() - ref_o).abs().max()))
        if lse is not None:
            err_l = max(err_l, float((lse[:, h0:h0 + 4, idx].double() - ref_l).abs().max()))
        del qs, s, ref_l, ref_o
    return err_o, err_l


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

    # the host/PCIe ceiling for the same bytes, all ranks copying at once (explains what is left of e2e at N > 1)
    if world > 1:
        dist.barrier()
    floor_ms = copy_floor_ms(torch, hosts, o_host, devs)
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, floor_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, floor_ms = t.tolist()
    want_multi = world > 1 and not args.no_extra
    if rank != 0:
        if want_multi:
            guarded_multi(None, args, torch, dist, fab, rank, world, device, flush)
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
                "api": "fa_forward_host (C-ABI, pinned host buffers)",
                "host_copy_floor_ms": round(floor_ms, 4),
                "host_copy_floor_note": "raw pinned-host<->device copies of the same bytes, no kernel, all ranks at once, max over ranks: "
                                        "what the host memory / PCIe path allows at this N; e2e minus this is the operator's own cost",
                "host_binding": HOST_BINDING},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s",
                     "frac": round(achieved / peak, 4), "traffic": TRAFFIC_BYTES.get(args.workload),
                     "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full "
                                       "capture (profiles/r02_ncu_summary.md), not re-measured in this run",
                     "peak_source": peaks["src"] + (" bf16_tflops / 2 (tf32, derived)" if dtype == "f32" else " bf16_tflops (burst)"),
                     "kernel": "fa_fwd_sm100_kernel", "hbm_gbs_achieved": round(bytes_of(B, H, N, d, es) / (ms_per_step * 1e-3) * 1e-9, 1)},
        "ms_min": round(min(ms), 5), "ms_median": round(sorted(ms)[len(ms) // 2], 5), "wall_s": round(wall, 3),
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(torch, N, d)
        line["cpu_baseline"]["reference_cpu_loop"] = reference_cpu_loop(5.0)   # the reference's own scalar CPU loop, beside it
    if world == 1 and not args.no_extra:
        line["other_configs"] = other_configs(torch, fab, device, flush, peaks)
    if want_multi:
        line.update(guarded_multi(line, args, torch, dist, fab, rank, world, device, flush))
    return line


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/); None = not captured
TRAFFIC_BYTES = {"C1": 12.66e6, "C2": 116.4e6, "C3": 52.6e6, "C4": 1058.5e6}   # dram__bytes_read+write per launch, profiles/r02_ncu_summary.md (kernel v10)


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
    # FA_FLAG_PRECISE (3xTF32) beside the default tf32 instance on the headline shape, and the llm.c harness shape through the
    # packed-QKV entry (what the exported attention_forward symbol runs: precise by default, tf32 on request)
    B, H, N, d, _ = WORKLOADS["C2"]
    _, (q, k, v) = make_inputs(torch, B, H, N, d, "f32", device, 99)
    out = torch.empty_like(q)
    ms = sorted(time_kernel(torch, lambda: fab.attention(q, k, v, scale=1.0 / math.sqrt(d), out=out, precise=True), 10, 3, flush))[5]
    res["C2_precise_3xtf32"] = {"ms": round(ms, 5), "tflops_algorithmic": round(flops_of(B, H, N, d) / ms * 1e-9, 1),
                                "note": "three tcgen05 MMAs per contraction: the tensor pipe does 3x these FLOPs"}
    del q, k, v, out
    Bl, Tl, Cl, NHl = 6, 4096, 768, 12
    inp = torch.rand(Bl, Tl, 3 * Cl, device=device) * 2 - 1
    outl = torch.empty(Bl, Tl, Cl, device=device)
    fl_l = 4.0 * Bl * NHl * (Cl // NHl) * Tl * (Tl + 1) / 2
    for label, prec in (("llmc_B6_T4096_precise", True), ("llmc_B6_T4096_tf32", False)):
        ms = sorted(time_kernel(torch, lambda: fab.attention_forward(6, outl, inp, Bl, Tl, Cl, NHl, 256, precise=prec), 10, 3, flush))[5]
        res[label] = {"ms": round(ms, 5), "tflops_causal": round(fl_l / ms * 1e-9, 1)}
    del inp, outl
    # the HBM-bound end of the operator: a decode-like launch (one query row per head, 131072 keys), split over the K/V axis
    # across CTAs and merged by the combine kernel; algorithmic bytes = K and V read once
    q = torch.randn(32, 1, 128, device=device).to(torch.bfloat16)
    k, v = (torch.randn(32, 131072, 128, device=device).to(torch.bfloat16) for _ in range(2))
    out = torch.empty_like(q)
    ms = sorted(time_kernel(torch, lambda: fab.attention(q, k, v, out=out), 10, 3, flush))[5]
    kv_bytes = 2.0 * k.numel() * 2
    res["decode_bh32_nq1_nk131072_bf16_d128"] = {"ms": round(ms, 5), "kv_gbs": round(kv_bytes / ms * 1e-6, 1),
                                                 "frac_hbm_peak": round(kv_bytes / ms * 1e-6 / peaks["hbm"], 4),
                                                 "note": "split-KV across CTAs + fa_combine_splits_kernel (2 launches); bytes = K + V read once"}
    del q, k, v, out
    # the backward pass (not in the reference, which is forward only) on the C4 shape: statistics + dK/dV + dQ launches.
    # FLOPs: algorithmic = 2.5 x forward (five contractions); executed = 3.5 x (the dQ launch recomputes S and dP instead of reducing
    # dQ across CTAs with atomics).  torch's scaled_dot_product_attention flash backward on the same tensors is timed beside it as a
    # library yardstick (a library kernel, not the reference).
    B, H, N, d, _ = WORKLOADS["C4"]
    for label, causal in (("C4_backward_bf16", False), ("C4_backward_bf16_causal", True)):
        g = torch.Generator(device=device).manual_seed(11)
        q, k, v, do = (torch.randn(B, H, N, d, device=device, generator=g).to(torch.bfloat16) for _ in range(4))
        scale = 1.0 / math.sqrt(d)
        o, lse = fab.attention(q, k, v, causal=causal, scale=scale, return_lse=True)
        ms = sorted(time_kernel(torch, lambda: fab.attention_backward(q, k, v, o, lse, do, causal=causal, scale=scale), 6, 2, flush))[3]
        alg = 2.5 * flops_of(B, H, N, d) * (0.5 if causal else 1.0)
        row = {"ms": round(ms, 4), "tflops_algorithmic_5_contractions": round(alg / ms * 1e-9, 1),
               "tflops_executed_7_contractions": round(1.4 * alg / ms * 1e-9, 1),
               "frac_tensor_peak_executed": round(1.4 * alg / ms * 1e-9 / peaks["bf16"], 4), "launches": 3}
        try:
            qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
            with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.FLASH_ATTENTION):
                oo = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=causal, scale=scale)
                row["torch_sdpa_flash_backward_ms"] = round(sorted(time_kernel(torch, lambda: oo.backward(do, retain_graph=True), 4, 1, flush))[2], 4)
            del qq, kk, vv, oo
        except Exception as ex:  # noqa: BLE001  (a yardstick only: its absence is reported, not fatal)
            row["torch_sdpa_flash_backward_ms"] = None
            row["torch_sdpa_note"] = repr(ex)[:100]
        res[label] = row
        del q, k, v, do, o, lse
    # config 5 on ONE GPU: the whole 131072-long sequence in one launch (the ring's single-GPU baseline)
    Hh, Nn, dd = 32, 131072, 128
    g = torch.Generator(device=device).manual_seed(5)
    q, k, v = (torch.randn(1, Hh, Nn, dd, device=device, generator=g).to(torch.bfloat16) for _ in range(3))
    out = torch.empty_like(q)
    ms = min(time_kernel(torch, lambda: fab.attention(q, k, v, out=out), 2, 1, flush))
    res["C5_one_gpu"] = {"ms": round(ms, 3), "tflops": round(flops_of(1, Hh, Nn, dd) / ms * 1e-9, 1),
                         "frac_tensor_peak_sustained": round(flops_of(1, Hh, Nn, dd) / ms * 1e-9 / peaks["bf16_sustained"], 4)}
    del q, k, v, out
    return res


def guarded_multi(line, args, torch, dist, fab, rank, world, device, flush, limit_s=300.0):
    """multi_gpu_sections under a deadline: if a rank dies or a collective hangs there, rank 0 still prints the headline line
    (already complete at this point) with the failure noted, and every rank exits instead of waiting for the NCCL timeout."""
    def bail():
        if line is not None:
            print(json.dumps({**line, "multi_gpu_error": f"multi-GPU sections did not finish within {limit_s:.0f} s"}), flush=True)
        os._exit(0)

    timer = threading.Timer(limit_s, bail)
    timer.daemon = True
    timer.start()
    try:
        return multi_gpu_sections(args, torch, dist, fab, rank, world, device, flush)
    finally:
        timer.cancel()


def multi_gpu_sections(args, torch, dist, fab, rank, world, device, flush):
    """The multi-GPU workloads BASELINE.json names, measured and parity-checked inside the driver's own `--gpus N` run:
    configs 4 and 3 at global size, B*H sharded (strong scaling, no collective), and config 5 as a ring."""
    res = {}
    peaks = load_peaks()
    for key, fn in (("c4_sharded", lambda: sharded_strong(torch, dist, fab, "C4", rank, world, device, flush, peaks)),
                    ("c3_sharded", lambda: sharded_strong(torch, dist, fab, "C3", rank, world, device, flush, peaks)),
                    ("c5_ring", lambda: ring_c5(torch, dist, fab, rank, world, device, peaks))):
        try:
            res[key] = fn()
        except Exception as e:  # a failure here must not take the headline line with it
            res[key] = {"error": repr(e)[:300]}
        torch.cuda.synchronize()
    return res


def sharded_strong(torch, dist, fab, name, rank, world, device, flush, peaks, steps=10):
    """Config `name` at its GLOBAL size, the flattened B*H axis cut into `world` contiguous slices (no data-path collective)."""
    B, H, N, d, dtype = WORKLOADS[name]
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    es = 2 if dtype == "bf16" else 4
    g = torch.Generator(device=device).manual_seed(4242)          # the same global tensors on every rank
    q, k, v = (torch.randn(B * H, N, d, device=device, generator=g).to(tdt) for _ in range(3))
    s0, s1 = fab.bh_shard_range(B * H, rank, world)
    qs, ks, vs = q[s0:s1], k[s0:s1], v[s0:s1]
    scale = 1.0 / math.sqrt(d)
    # parity (1): this rank's shard == the same rows of the unsharded forward, bit for bit (batch-invariant scheduling)
    o_full = fab.attention(q, k, v, scale=scale, batch_invariant=True)
    bit_equal = bool(torch.equal(fab.attention(qs, ks, vs, scale=scale, batch_invariant=True), o_full[s0:s1]))
    # parity (2): default scheduling, 64 sampled rows per head against fp64 over all keys
    o_sh, lse_sh = fab.attention(qs, ks, vs, scale=scale, return_lse=True)
    err_o, err_l = fp64_rows_check(torch, qs[None], ks[None], vs[None], o_sh[None], lse_sh[None], scale, 64, 7 + rank)
    out_n, out_1 = torch.empty_like(qs), torch.empty_like(q)
    ms_n = time_kernel(torch, lambda: fab.attention(qs, ks, vs, scale=scale, out=out_n), steps, 3, flush)
    ms_1 = time_kernel(torch, lambda: fab.attention(q, k, v, scale=scale, out=out_1), steps, 3, flush)
    med = lambda xs: sorted(xs)[len(xs) // 2]      # noqa: E731  (a 20-50 us kernel: one slow step on one of N ranks would own a mean)
    t_max = torch.tensor([med(ms_n), err_o, err_l, 0.0 if bit_equal else 1.0], dtype=torch.float64, device=device)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    t_sum = torch.tensor([med(ms_1)], dtype=torch.float64, device=device)
    dist.all_reduce(t_sum)
    t_n, err_o, err_l, not_equal = t_max.tolist()
    t_1 = float(t_sum[0]) / world
    fl = flops_of(B, H, N, d)
    peak = peaks["bf16"] if dtype == "bf16" else peaks["bf16"] / 2
    tol = 2e-2 if dtype == "bf16" else 1e-3
    return {"workload": f"{name}: B={B} H={H} d={d} N={N} {dtype}, global B*H={B * H} cut into {world} slices of {s1 - s0}, no collective",
            "scaling": "strong", "ms": round(t_n, 5), "tflops_total": round(fl / t_n * 1e-9, 1),
            "frac_tensor_peak_per_gpu": round(fl / world / t_n * 1e-9 / peak, 4),
            "hbm_gbs_per_gpu": round(bytes_of(B, H, N, d, es) / world / t_n * 1e-6, 1),
            "ms_one_gpu_same_run": round(t_1, 5), "speedup_vs_one_gpu": round(t_1 / t_n, 3), "efficiency": round(t_1 / (world * t_n), 4),
            "timing": "CUDA events, L2 flushed before every step, median of 10 steps, max over ranks",
            "parity": {"shard_bit_equal_to_unsharded_forward": not_equal == 0.0, "max_abs_err_o_vs_fp64_sampled_rows": err_o,
                       "max_abs_err_lse_vs_fp64_sampled_rows": err_l, "rows_per_head": 64, "tolerance_o": tol,
                       "ok": bool(not_equal == 0.0 and err_o < tol)}}


def ring_c5(torch, dist, fab, rank, world, device, peaks, steps=3):
    """Config 5: B=1 H=32 d=128 N=131072 bf16, the sequence cut into `world` shards; K/V shards pulled from their owners over
    NVLink by the copy engines while the kernel of the current step runs; every step's kernel merges into the running
    (O, LSE) in its epilogue."""
    H, N, d = 32, 131072, 128
    n_loc = N // world
    scale = 1.0 / math.sqrt(d)
    g = torch.Generator(device=device).manual_seed(5151)          # the same full tensors on every rank (1.07 GB each)
    qf, kf, vf = (torch.randn(1, H, N, d, device=device, generator=g).to(torch.bfloat16) for _ in range(3))
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    q, k, v = (t[:, :, sl].contiguous() for t in (qf, kf, vf))
    del qf

    def timed(f, reps, warm):
        for _ in range(warm):
            f()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    transport = "p2p"
    try:
        fab.ring_attention(q, k, v, transport="p2p")
    except fab.FaError as e:        # raised on every rank alike (the transport's set-up is agreed between the ranks)
        transport = "nccl"
        transport_note = f"p2p unavailable ({str(e)[:120]}): NCCL send/recv rotation"
    else:
        transport_note = "p2p (copy-engine pulls from CUDA-IPC-mapped peer buffers over NVLink)"
    def timed_local(f, reps):     # no collective inside: for work only one rank does
        tot = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    launches0 = fab.launch_count()
    o, lse = fab.ring_attention(q, k, v, transport=transport)
    launches = fab.launch_count() - launches0
    # parity (1): this rank's ring result vs ONE kernel launch of its queries over the full key sequence
    o_one, lse_one = fab.attention(q, kf, vf, scale=scale, return_lse=True)
    d_o = float((o.float() - o_one.float()).abs().max())
    d_l = float((lse - lse_one).abs().max())
    # parity (2): 16 sampled rows per head against fp64 over all 131072 keys
    err_o, err_l = fp64_rows_check(torch, q, kf, vf, o, lse, scale, 16, 100 + rank)
    del o_one, lse_one
    # the same problem on ONE GPU (one launch over the whole sequence), twice: with every GPU of the box running it at the same
    # time (the box's power / clock state is then that of the ring run) and on rank 0 alone with the others idle
    qf = torch.cat([q] * world, dim=2)[:, :, :N]          # a full-length query tensor (the values do not matter for timing)
    out_full = torch.empty_like(qf)

    def one_gpu():
        fab.attention(qf, kf, vf, scale=scale, out=out_full)

    ms_one_busy = timed(one_gpu, 1, 1)
    ms_one_alone = 0.0
    dist.barrier()
    if rank == 0:
        ms_one_alone = timed_local(one_gpu, 1)
    dist.barrier()
    del qf, out_full, kf, vf
    sampler = ClockSampler(torch.cuda.current_device())     # SM clock / power-cap state while all ranks run the ring
    sampler.start()
    ms_ring = timed(lambda: fab.ring_attention(q, k, v, transport=transport), steps, 1)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    ms_nccl = timed(lambda: fab.ring_attention(q, k, v, transport="nccl"), max(1, steps - 1), 1)

    def local_only():   # the same kernels (merge fused in the epilogue) on the resident shard, no transfers
        acc = list(fab.attention(q, k, v, scale=scale, return_lse=True, out_f32=True))
        for s_ in range(1, world):
            acc[0] = fab.attention(q, k, v, scale=scale, out_f32=s_ < world - 1, acc=(acc[0], acc[1]))

    ms_local = timed(local_only, steps, 1)
    t = torch.tensor([ms_ring, ms_nccl, ms_local, d_o, d_l, err_o, err_l, ms_one_busy, ms_one_alone], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_ring, ms_nccl, ms_local, d_o, d_l, err_o, err_l, ms_one_busy, ms_one_alone = t.tolist()
    fl = 4.0 * H * float(N) * N * d
    per_gpu = fl / world / ms_ring * 1e-9
    return {"workload": f"C5: B=1 H={H} d={d} N={N} bf16 non-causal, sequence cut into {world} shards of {n_loc}",
            "scaling": "strong", "transport": transport_note,
            "ms": round(ms_ring, 3), "tflops_total": round(fl / ms_ring * 1e-9, 1), "tflops_per_gpu": round(per_gpu, 1),
            "frac_sustained_peak": round(per_gpu / peaks["bf16_sustained"], 4), "frac_burst_peak": round(per_gpu / peaks["bf16"], 4),
            "ms_same_kernels_no_transfers": round(ms_local, 3), "overlap": round(ms_local / ms_ring, 4),
            "ms_one_gpu_same_run_all_gpus_busy": round(ms_one_busy, 3), "speedup_vs_one_gpu_all_busy": round(ms_one_busy / ms_ring, 3),
            "ms_one_gpu_same_run_alone": round(ms_one_alone, 3), "speedup_vs_one_gpu_alone": round(ms_one_alone / ms_ring, 3),
            "ms_nccl_transport": round(ms_nccl, 3), "kernel_launches_per_forward": launches,
            "kv_bytes_pulled_per_gpu": 2 * k.numel() * 2 * (world - 1),
            "timing": f"CUDA events around one ring forward, barrier before each, mean of {steps}, max over ranks",
            "clocks_during_ring_rank0": sampler.summary(),
            "parity": {"max_abs_diff_o_vs_one_kernel_over_full_sequence": d_o, "max_abs_diff_lse_vs_one_kernel": d_l,
                       "max_abs_err_o_vs_fp64_sampled_rows": err_o, "max_abs_err_lse_vs_fp64_sampled_rows": err_l, "rows_per_head": 16,
                       "tolerance_o": 2e-2, "ok": bool(d_o < 2e-2 and err_o < 2e-2 and err_l < 2e-3)}}


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
                     "kernel": "fa_fwd_sm100_kernel (bf16 d=128) x ring steps, each merging into the running (O, LSE) in its epilogue"},
    }


def run_reference(args, torch, dist, rank, world, device):
    """The reference arm: its own forward(Q,K,V,causal) (scale fixed at 1.0 inside, src/flashattention.cu:593).  Its
    implementation of this path is a CUDA kernel, so at N > 1 EVERY rank runs it on its own GPU on the same per-GPU workload as
    our arm (weak scaling, max over ranks): the driver's same-N ratio then compares N GPUs with N GPUs."""
    from oracle import fa_oracle

    long_seq = args.workload == "C5"     # one forward of the reference kernel takes ~15 s at N = 131072: one warm-up, one step
    B, H, N, d, dtype = (1, 32, 131072, 128, "bf16") if long_seq else WORKLOADS[args.workload]
    fl = flops_of(B, H, N, d)
    base = {"metric": "fwd attention TFLOP/s (B2 H8 d64 N8192)" if args.workload == "C2" else f"fwd attention TFLOP/s ({args.workload})",
            "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic", "impl": "reference"}
    ext = fa_oracle.load_ref_torch_ext(d)  # bf16 workloads: the reference runs fp32 copies, it has no bf16 path
    if ext is not None and torch.cuda.is_available():
        g = torch.Generator().manual_seed(1234 + rank)
        hosts = [torch.randn(B * H, N, d, generator=g).pin_memory() for _ in range(3)]
        q, k, v = (h.to(device) for h in hosts)
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
        steps = 1 if long_seq else max(1, args.steps)          # one forward takes ~15 s at N = 131072
        warm = 1 if long_seq else max(1, args.warmup)
        if world > 1:
            dist.barrier()
        ms = time_kernel(torch, lambda: ext.forward(q, k, v, False), steps, warm, flush)
        ms_per_step = sum(ms) / len(ms)
        o_host = torch.empty(B * H, N, d).pin_memory()
        e2e_steps = 1 if long_seq else max(1, min(steps, 10))
        if world > 1:
            dist.barrier()
        e2e_each = []
        for it in range(e2e_steps + 1):      # one untimed pass first (allocator, page registration), like our own arm's
            t0 = time.perf_counter()
            dq, dk, dv = (h.to(device, non_blocking=True) for h in hosts)
            o_host.copy_(ext.forward(dq, dk, dv, False))
            torch.cuda.synchronize()
            if it:
                e2e_each.append(time.perf_counter() - t0)
        e2e_s = sum(e2e_each) / len(e2e_each)
        if world > 1:
            t = torch.tensor([ms_per_step, e2e_s], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_per_step, e2e_s = t.tolist()
        if rank != 0:
            return None
        fl = fl * world
        base.update({"value": round(fl / (ms_per_step * 1e-3) * 1e-12, 3), "ms_per_step": round(ms_per_step, 4), "steps": steps, "warmup": warm,
                     "dtype": "f32", "config": {"workload": f"{args.workload}: B={B} H={H} d={d} N={N} fp32 non-causal; reference CUDA kernel "
                                                            "flash_tiled_coarse rebuilt for sm_100a (oracle/_ref), its forward() incl. torch::zeros + cudaDeviceSynchronize, scale fixed at 1.0"},
                     "e2e": {"value": round(fl / e2e_s * 1e-12, 3), "unit": "TFLOP/s", "ms_per_step": round(e2e_s * 1e3, 3),
                             "h2d_bytes_per_step": 3 * q.numel() * 4, "d2h_bytes_per_step": q.numel() * 4, "steps": e2e_steps,
                             "ms_min_rank0": round(min(e2e_each) * 1e3, 3), "ms_median_rank0": round(sorted(e2e_each)[len(e2e_each) // 2] * 1e3, 3),
                             "host_binding": HOST_BINDING},
                     "cpu_baseline": reference_cpu_loop() or {
                         "value": None, "unit": "TFLOP/s", "cores": 0, "kind": "reference",
                         "sample": "the reference for this path is a CUDA kernel; it ran on the GPU, not on host cores"}})
        base["cpu_baseline"]["note"] = ("the reference's implementation of this path is a CUDA kernel: `value` and `e2e` of this line are "
                                        "that kernel on the same GPU; cpu_baseline is the only CPU code the reference has for it")
        return base
    # no reference build on this box: time the oracle's CPU port of the same algorithm on a bounded sample (rank 0 only)
    if rank != 0:
        return None
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
    if not torch.cuda.is_available():
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps(run_reference(args, torch, dist, 0, world, None)))
            return 0
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    # cores / NUMA node first: the pinned host buffers allocated below land where this process runs
    HOST_BINDING.clear()
    HOST_BINDING.update(pin_to_gpu_numa(local, world))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import datetime

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=300))
    try:
        if args.impl == "ours" and args.workload == "C5":
            line = run_ring(args, torch, dist, rank, world, device)
        elif args.impl == "ours":
            line = run_ours(args, torch, dist, rank, world, device)
        else:
            line = run_reference(args, torch, dist, rank, world, device)
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1 and dist.is_initialized():
            if args.impl == "ours":
                try:
                    import flashattention_c_b200 as fab

                    fab.ring_p2p_release()      # collective: unmap peers' K/V buffers before anyone frees its own
                except Exception:
                    pass
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
