"""A few forwards at a given shape (profiling target): python scripts/fwd_one.py B H N d dtype(f32|bf16|f16) causal [reps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import flashattention_c_b200 as fab  # noqa: E402

B, H, N, d = (int(x) for x in sys.argv[1:5])
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[5]]
causal = bool(int(sys.argv[6]))
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
q, k, v = (torch.randn(B, H, N, d, generator=g, device=dev).to(dt) for _ in range(3))
out = torch.empty_like(q)
for _ in range(reps):
    fab.attention(q, k, v, causal=causal, out=out)
torch.cuda.synchronize()
