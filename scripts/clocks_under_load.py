"""SM clock / power while the forward kernel runs back to back (NVML samples every 10 ms), per workload.
Answers: is a config power-limited (clock below max under load)?   python scripts/clocks_under_load.py"""
import sys, time, threading, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, pynvml
import flashattention_c_b200 as fab

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
W = {"C2": (16, 8192, 64, torch.float32), "C4": (128, 8192, 128, torch.bfloat16), "C3": (128, 1024, 32, torch.float32)}
for name, (bh, n, d, dt) in W.items():
    q, k, v = (torch.randn(1, bh, n, d, device="cuda").to(dt) for _ in range(3))
    for _ in range(3): fab.attention(q, k, v)
    torch.cuda.synchronize()
    samples, stop = [], False
    def sample():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(0.01)
    th = threading.Thread(target=sample); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(50, int(2.0 / {"C2": 0.0004, "C4": 0.0033, "C3": 0.00006}[name]))
    t0 = time.time(); e0.record()
    for _ in range(reps): fab.attention(q, k, v)
    e1.record(); torch.cuda.synchronize()
    stop = True; th.join()
    ms = e0.elapsed_time(e1) / reps
    clk = sorted(s[0] for s in samples[len(samples) // 4:]); pw = sorted(s[1] for s in samples[len(samples) // 4:])
    reasons = 0
    for s in samples: reasons |= s[2]
    flops = 4.0 * bh * n * n * d
    print(json.dumps({"workload": name, "reps": reps, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "sm_mhz_median": clk[len(clk) // 2],
                      "sm_mhz_min": clk[0], "power_w_median": pw[len(pw) // 2], "power_w_max": pw[-1], "throttle_reasons_mask": hex(reasons),
                      "samples": len(samples)}), flush=True)
