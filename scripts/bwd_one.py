"""One forward + one backward at a given shape (profiling target): python scripts/bwd_one.py B H N d causal [reps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import flashattention_c_b200 as fab  # noqa: E402

B, H, N, d, causal = (int(x) for x in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
q, k, v, do = (torch.randn(B, H, N, d, generator=g, device=dev).to(torch.bfloat16) for _ in range(4))
o, lse = fab.attention(q, k, v, causal=bool(causal), return_lse=True)
for _ in range(reps):
    fab.attention_backward(q, k, v, o, lse, do, causal=bool(causal))
torch.cuda.synchronize()
