#!/usr/bin/env python
"""Reference CUDA kernel (oracle/_ref/flash_ref_d{32,64,128}.so = src/main.cpp + src/flashattention.cu rebuilt for sm_100a)
vs this repo on every BASELINE shape, same B200, same run: kernel-only CUDA-event times.  Both kernels get IDENTICAL values:
for the bf16 configs the draw is rounded to bf16 first and the reference (which has no bf16 path) runs on the fp32 copy of
the rounded tensors, so `max_abs_diff_vs_reference_kernel_scale1` measures the kernels, not an input rounding.  The fp32
configs also report the precise (3xTF32) instance where it exists (d <= 64)."""
import json
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import flashattention_c_b200 as fab  # noqa: E402
from oracle import fa_oracle  # noqa: E402

SHAPES = [("C1", 16, 1024, 64, torch.float32), ("C2", 16, 8192, 64, torch.float32), ("C3", 128, 1024, 32, torch.float32),
          ("C4", 128, 8192, 128, torch.bfloat16), ("C5 (one GPU, full N)", 32, 131072, 128, torch.bfloat16)]


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


for name, bh, n, d, dt in SHAPES:
    big = n > 10000
    g = torch.Generator(device="cuda").manual_seed(1)
    qq, kk, vv = (torch.randn(bh, n, d, device="cuda", generator=g).to(dt) for _ in range(3))
    q, k, v = (x.float() for x in (qq, kk, vv))     # what the reference reads: the same values, as fp32
    ext = fa_oracle.load_ref_torch_ext(d)
    t_ref = timeit(lambda: ext.forward(q, k, v, False), 1 if big else 3)
    o_ref = ext.forward(q, k, v, False)
    t_ours = timeit(lambda: fab.attention(qq, kk, vv, scale=1.0), 3 if big else 10)
    o = fab.attention(qq, kk, vv, scale=1.0).float()
    fl = 4.0 * bh * n * n * d
    extra = {}
    if dt == torch.float32 and d <= 64:
        t_p = timeit(lambda: fab.attention(qq, kk, vv, scale=1.0, precise=True), 10)
        o_p = fab.attention(qq, kk, vv, scale=1.0, precise=True)
        extra = {"precise_ms": round(t_p, 4), "precise_max_abs_diff_vs_reference_kernel_scale1": float((o_p - o_ref).abs().max())}
    print(json.dumps({"config": name, "bh": bh, "n": n, "d": d, "dtype": str(dt).split(".")[-1], "ref_ms": round(t_ref, 3),
                      "ref_tflops": round(fl / t_ref * 1e-9, 2), "ours_ms": round(t_ours, 4), "ours_tflops": round(fl / t_ours * 1e-9, 1),
                      "speedup": round(t_ref / t_ours, 1), "max_abs_diff_vs_reference_kernel_scale1": float((o - o_ref).abs().max()),
                      # relative to the output's magnitude: a 16-bit O carries one ulp of its own format (bf16: 2^-8 relative, 0.031 at |O| in [4, 8))
                      "max_rel_diff_vs_reference_kernel_scale1": float(((o - o_ref).abs() / o_ref.abs().clamp_min(1.0)).max()), **extra}), flush=True)
    del q, k, v, qq, kk, vv, o, o_ref


# ---- the llm.c surface: the reference's own attention_forward6 (permute -> flashattention -> unpermute, cudaMalloc/cudaFree and four
# device synchronisations inside: src/llm.c/attention_forward.cu:1106-1179) vs the exported symbol of this repo, harness shape
import contextlib
import ctypes
import os

ref6 = fa_oracle.ref_llmc_gpu_entry()
if ref6 is not None:
    L = fab.lib()
    L.attention_forward6.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5
    L.attention_forward6.restype = None
    B, T, C, NH = 6, 4096, 768, 12
    inp = torch.rand(B, T, 3 * C, device="cuda") * 2 - 1
    o_ref, o_p, o_t = (torch.empty(B, T, C, device="cuda") for _ in range(3))
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(devnull.fileno(), 1)       # the reference prints a timing line per call
        try:
            t_ref = timeit(lambda: ref6(o_ref.data_ptr(), inp.data_ptr(), B, T, C, NH, 256), 5)
        finally:
            os.dup2(saved, 1)
    t_p = timeit(lambda: L.attention_forward6(o_p.data_ptr(), inp.data_ptr(), B, T, C, NH, 256), 10)
    t_t = timeit(lambda: fab.attention_forward(6, o_t, inp, B, T, C, NH, 256, precise=False), 10)
    print(json.dumps({"config": "llm.c harness shape B6 T4096 C768 NH12 (causal, packed QKV), attention_forward6", "ref_ms": round(t_ref, 3),
                      "ours_precise_ms": round(t_p, 4), "ours_tf32_ms": round(t_t, 4), "speedup_precise": round(t_ref / t_p, 1),
                      "max_abs_diff_precise_vs_reference_entry": float((o_p - o_ref).abs().max()),
                      "max_abs_diff_tf32_vs_reference_entry": float((o_t - o_ref).abs().max())}), flush=True)
