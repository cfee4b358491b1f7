import os, sys, time, math, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
if len(sys.argv) > 1:
    import torch, flashattention_c_b200 as fab
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(2, 8, 8192, 64, generator=g).pin_memory() for _ in range(3))
    o = torch.empty_like(q).pin_memory()
    for _ in range(3): fab.attention_host(q, k, v, out=o)
    t0 = time.perf_counter()
    for _ in range(10): fab.attention_host(q, k, v, out=o)
    print(f"chunks={sys.argv[1]}: e2e {(time.perf_counter()-t0)/10*1e3:.3f} ms")
else:
    for c in (1, 2, 4, 8, 16):
        e = dict(os.environ); e["FA_B200_HOST_CHUNKS"] = str(c)
        subprocess.run([sys.executable, __file__, str(c)], env=e)
