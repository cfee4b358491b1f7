"""e2e (host buffers -> O in host buffers) timing of fa_forward_host under different chunk schedules, plus the raw
pinned-memory copy rates of the box that bound it.  Run on a GPU box: python scripts/e2e_chunks.py"""
import os, sys, time, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
if len(sys.argv) > 1:
    import torch, flashattention_c_b200 as fab
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(2, 8, 8192, 64, generator=g).pin_memory() for _ in range(3))
    o = torch.empty_like(q).pin_memory()
    for _ in range(3): fab.attention_host(q, k, v, out=o)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); fab.attention_host(q, k, v, out=o); ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"{sys.argv[1]}: e2e median {ts[len(ts)//2]*1e3:.3f} ms  min {ts[0]*1e3:.3f} ms", flush=True)
    if sys.argv[1] == "raw":
        d = [torch.empty_like(q, device="cuda") for _ in range(4)]
        def timed(fn, n=10):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n): fn()
            torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
        t = timed(lambda: [d[i].copy_(x, non_blocking=True) for i, x in enumerate((q, k, v))])
        print(f"raw H2D 3x33.5MB: {t*1e3:.3f} ms = {100.66e6/t/1e9:.1f} GB/s")
        t = timed(lambda: o.copy_(d[3], non_blocking=True))
        print(f"raw D2H 33.5MB: {t*1e3:.3f} ms = {33.55e6/t/1e9:.1f} GB/s")
        s2 = torch.cuda.Stream()
        def both():
            [d[i].copy_(x, non_blocking=True) for i, x in enumerate((q, k, v))]
            with torch.cuda.stream(s2): o.copy_(d[3], non_blocking=True)
        t = timed(both)
        print(f"raw H2D 100MB || D2H 33.5MB: {t*1e3:.3f} ms")
else:
    runs = [("raw", {}), ("default(decay .4)", {}), ("decay .3", {"FA_B200_HOST_DECAY": ".3"}), ("decay .5", {"FA_B200_HOST_DECAY": ".5"}),
            ("decay .6", {"FA_B200_HOST_DECAY": ".6"})]
    runs += [(f"equal x{c}", {"FA_B200_HOST_CHUNKS": str(c)}) for c in (1, 2, 4, 8)]
    runs += [(f"sched {s}", {"FA_B200_HOST_SCHED": s}) for s in ("6,5,3,1,1", "8,4,2,1,1", "4,4,4,2,1,1", "12,3,1", "5,4,3,2,1,1")]
    for name, env in runs:
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, name], env=e)
