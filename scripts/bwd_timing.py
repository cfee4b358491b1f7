"""Timing of the backward kernels (CUDA events on the launch stream, L2 flushed between iterations): total and per launch.
FLOPs: algorithmic = 10 B H N_q N_k d (five contractions: the single-pass count; causal halves it); executed = 14 B H N_q N_k d
(seven: S and dP are recomputed by the dQ launch)."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import flashattention_c_b200 as fab  # noqa: E402

CONFIGS = [
    ("C4_bwd_B4_H32_N8192_d128_bf16", 4, 32, 8192, 128, torch.bfloat16, False),
    ("C4_bwd_causal", 4, 32, 8192, 128, torch.bfloat16, True),
    ("B2_H8_N8192_d64_bf16", 2, 8, 8192, 64, torch.bfloat16, False),
    ("B8_H16_N2048_d64_bf16_causal", 8, 16, 2048, 64, torch.bfloat16, True),
    ("B1_H32_N32768_d128_bf16_causal", 1, 32, 32768, 128, torch.bfloat16, True),
]


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    for name, B, H, N, d, dt, causal in CONFIGS:
        g = torch.Generator(device=dev).manual_seed(1)
        q, k, v, do = (torch.randn(B, H, N, d, generator=g, device=dev).to(dt) for _ in range(4))
        scale = d ** -0.5
        o, lse = fab.attention(q, k, v, causal=causal, scale=scale, return_lse=True)
        ms_f, ms_b = [], []
        for it in range(steps + 2):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            fab.attention(q, k, v, causal=causal, scale=scale, return_lse=True)
            e[1].record()
            fab.attention_backward(q, k, v, o, lse, do, causal=causal, scale=scale)
            e[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                ms_f.append(e[0].elapsed_time(e[1]))
                ms_b.append(e[1].elapsed_time(e[2]))
        mf, mb = sorted(ms_f)[len(ms_f) // 2], sorted(ms_b)[len(ms_b) // 2]
        alg = 10.0 * B * H * N * N * d * (0.5 if causal else 1.0)
        line = {"config": name, "fwd_ms": round(mf, 4), "bwd_ms": round(mb, 4), "bwd_over_fwd": round(mb / mf, 3),
                "bwd_tflops_algorithmic_5gemm": round(alg / mb / 1e9, 1), "bwd_tflops_executed_7gemm": round(1.4 * alg / mb / 1e9, 1)}
        # library baseline beside it: torch SDPA (flash backend) backward on the same tensors
        try:
            if os.environ.get("BWD_NO_TORCH"):
                raise RuntimeError("skipped (BWD_NO_TORCH)")
            qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
            with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.FLASH_ATTENTION):
                oo = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=causal, scale=scale)
                ts = []
                for it in range(steps + 2):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    oo.backward(do, retain_graph=True)
                    b.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(a.elapsed_time(b))
            line["torch_sdpa_flash_bwd_ms"] = round(sorted(ts)[len(ts) // 2], 4)
        except Exception as ex:  # noqa: BLE001
            line["torch_sdpa_flash_bwd_ms"] = None
            line["torch_sdpa_note"] = repr(ex)[:120]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
