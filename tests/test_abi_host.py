"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/fa_b200.h declares,
argument validation works without a device, and the host-side logic (sharding, ring schedule, error behaviour)."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "fa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|int64_t|char\s*\*|const char\s*\*)\s*\*?\s*(\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_symbols_match_python_binding(fab):
    from flashattention_c_b200 import _lib

    declared = _declared_symbols()
    assert declared, "could not parse include/fa_b200.h"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_library_loads_and_exports_every_declared_symbol(fab):
    L = fab.lib()
    for name in _declared_symbols():
        assert hasattr(L, name), f"libfa_b200.so does not export {name}"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "flashattention.c_b200" / "libfa_b200.so")],
                        capture_output=True, text=True, check=True).stdout
    for name in _declared_symbols():
        assert re.search(rf"\bT {name}\b", nm), f"{name} is not an unmangled (extern \"C\") export"


def test_ctypes_struct_layout_matches_the_header(tmp_path):
    """fa_params as gcc lays it out from include/fa_b200.h == the ctypes mirror (sizes and every field offset), and the
    flag / enum values the Python side hard-codes are the header's."""
    import ctypes
    import subprocess

    sys.path.insert(0, str(ROOT))
    from flashattention_c_b200 import _lib

    fields = [f[0] for f in _lib.FaParams._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "fa_b200.h"}"', 'int main(void) {',
            '  printf("%zu\\n", sizeof(fa_params));']
    prog += [f'  printf("%zu\\n", offsetof(fa_params, {f}));' for f in fields]
    prog += ['  printf("%d %d %d %d %d %d\\n", FA_F32, FA_BF16, FA_IMPL_TCGEN05, FA_IMPL_SIMT, FA_FLAG_BATCH_INVARIANT, FA_B200_VERSION);',
             '  return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.FaParams)
    for f, off in zip(fields, out[1:1 + len(fields)]):
        assert int(off) == getattr(_lib.FaParams, f).offset, f
    consts = [int(x) for x in out[1 + len(fields):]]
    assert consts[:5] == [_lib.FA_F32, _lib.FA_BF16, _lib.FA_IMPL_TCGEN05, _lib.FA_IMPL_SIMT, _lib.FA_FLAG_BATCH_INVARIANT]


def test_tma_view_passes_aligned_strided_views_and_copies_the_rest():
    """Host logic of api.attention: which views go to the kernel as they are (strides into the TMA map) and which are copied."""
    import torch

    sys.path.insert(0, str(ROOT))
    from flashattention_c_b200 import api

    x = torch.zeros(2, 3, 64, 32)
    v = x[:, :, 16:48]                      # sequence slice: last axis contiguous, strides multiples of 16 bytes
    assert api._tma_view(v) is v and api._strides_bhn(v) == (3 * 64 * 32, 64 * 32, 32)
    assert api._tma_view(x.transpose(2, 3)).is_contiguous()          # last axis strided -> copy
    assert api._tma_view(x[..., 1:31]).is_contiguous() and api._tma_view(x[..., 1:31]) is not x   # misaligned base -> copy
    y3 = torch.zeros(6, 64, 32)[:, 8:40]
    assert api._strides_bhn(y3) == (6 * 64 * 32, 64 * 32, 32)


def test_version_and_strerror(fab):
    L = fab.lib()
    assert L.fa_version() == 102
    assert L.fa_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.fa_strerror(-2)
    assert L.fa_strerror(-12345) == b"unknown status"


def test_kernel_dispatch_table(fab):
    """fa_query_instance (host only): fp32 d <= 32 / 64 / 128 and 16-bit d <= 64 / 128 / 256 map onto the tcgen05 instances
    (smaller head dims zero-padded by TMA), fp32 (128, 256] goes to the CUDA-core kernel, the rest is refused."""
    from flashattention_c_b200 import _lib

    q = fab.lib().fa_query_instance
    F32, BF16, F16 = _lib.FA_F32, _lib.FA_BF16, _lib.FA_F16
    assert [q(F32, d) for d in (4, 8, 32, 36, 64, 68, 96, 128)] == [32, 32, 32, 64, 64, 128, 128, 128]
    assert [q(F32, d) for d in (136, 256)] == [0, 0]                       # CUDA-core kernel
    assert q(F32, 6) == -4 and q(F32, 132) == -4 and q(F32, 264) == -4      # rows not 16-byte aligned / too wide
    for dt in (BF16, F16):
        assert [q(dt, d) for d in (8, 64, 72, 128, 136, 256)] == [64, 64, 128, 128, 256, 256]
        assert q(dt, 12) == -4 and q(dt, 264) == -4
    assert q(7, 64) == -1 and q(F32, 0) == -1


def test_backward_struct_layout_matches_the_header(tmp_path):
    """fa_bwd_params: gcc's layout of the header == the ctypes mirror."""
    sys.path.insert(0, str(ROOT))
    from flashattention_c_b200 import _lib

    fields = [f[0] for f in _lib.FaBwdParams._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "fa_b200.h"}"', 'int main(void) {',
            '  printf("%zu\\n", sizeof(fa_bwd_params));']
    prog += [f'  printf("%zu\\n", offsetof(fa_bwd_params, {f}));' for f in fields]
    prog += ['  return 0; }']
    src = tmp_path / "layout_bwd.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout_bwd"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.FaBwdParams)
    for f, off in zip(fields, out[1:]):
        assert int(off) == getattr(_lib.FaBwdParams, f).offset, f


def test_backward_arguments_are_rejected_before_touching_a_device(fab):
    from flashattention_c_b200 import _lib
    L = fab.lib()
    buf = (ctypes.c_float * 64)()
    ptr = ctypes.cast(buf, ctypes.c_void_p).value
    ptr = (ptr + 15) & ~15

    def params(**kw):
        p = _lib.FaBwdParams()
        for n in ("q", "k", "v", "o", "d_o", "lse", "dq", "dk", "dv"):
            setattr(p, n, ptr)
        p.batch, p.heads, p.kv_heads, p.n_q, p.n_k, p.head_dim, p.dtype, p.causal, p.scale = 1, 2, 0, 16, 16, 64, _lib.FA_BF16, 0, 1.0
        for t in ("q", "k", "v", "o", "do", "dq", "dk", "dv"):
            setattr(p, f"{t}_stride_b", 2 * 16 * 64)
            setattr(p, f"{t}_stride_h", 16 * 64)
            setattr(p, f"{t}_stride_n", 64)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    assert L.fa_backward(None, None) == -1
    assert L.fa_backward(ctypes.byref(params(dq=None)), None) == -1          # null gradient pointer
    assert L.fa_backward(ctypes.byref(params(n_k=0)), None) == -1
    assert L.fa_backward(ctypes.byref(params(scale=0.0)), None) == -1
    assert L.fa_backward(ctypes.byref(params(kv_heads=3)), None) == -1       # heads % kv_heads
    assert L.fa_backward(ctypes.byref(params(dtype=_lib.FA_F32)), None) == -4   # fp32: no backward instance (unsupported, not invalid)
    assert L.fa_backward(ctypes.byref(params(head_dim=256)), None) == -4
    assert L.fa_backward(ctypes.byref(params(dq=ptr + 4)), None) == -5       # misaligned gradient pointer
    assert L.fa_backward(ctypes.byref(params(dk_stride_n=68)), None) == -5   # gradient stride not a multiple of 16 bytes
    if not torch.cuda.is_available():
        assert L.fa_backward(ctypes.byref(params()), None) == -2            # FA_ERR_NO_DEVICE: no fallback


def test_invalid_arguments_are_rejected_before_touching_a_device(fab):
    L = fab.lib()
    assert L.fa_forward(None, None, None, None, None, 1, 1, 16, 16, 64, 1.0, 0, 0, None) == -1   # null pointers
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.fa_forward(p, p, p, p, None, 1, 1, 0, 16, 64, 1.0, 0, 0, None) == -1                # n_q = 0
    assert L.fa_forward(p, p, p, p, None, 1, 1, 16, 16, 64, 1.0, 0, 7, None) == -1               # bad dtype
    assert L.fa_forward(p, p, p, p, None, 1, 1, 16, 16, 64, -1.0, 0, 0, None) == -1              # scale <= 0
    assert L.fa_merge_partials(p, p, p, p, 4, 6, None) == -1                                     # head_dim % 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_device_is_an_error_not_a_fallback(fab):
    L = fab.lib()
    buf = (ctypes.c_float * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.fa_forward(p, p, p, p, None, 1, 1, 16, 16, 64, 1.0, 0, 0, None) == -2  # FA_ERR_NO_DEVICE
    q = torch.zeros(2, 16, 64)
    with pytest.raises(fab.FaError):
        fab.forward(q, q, q, False)   # CPU tensors: the operator has no CPU path
    with pytest.raises(fab.FaError):
        fab.attention_host(q, q, q)    # host entry needs a device too


def test_python_argument_checks(fab):
    q = torch.zeros(2, 16, 64)
    with pytest.raises(fab.FaError):
        fab.attention(q, q[:, :, :32], q)          # shape mismatch / CPU tensors
    with pytest.raises(fab.FaError):
        fab.attention_forward(3, q, q, 1, 16, 64, 1)  # kernels 1-5 of the llm.c harness are out of scope


def test_bh_shard_range_covers_everything_once(fab):
    for total in (1, 7, 16, 128, 129):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s, e = fab.bh_shard_range(total, r, world)
                assert 0 <= s <= e <= total
                seen += list(range(s, e))
            assert seen == list(range(total))


def test_ring_schedule_visits_every_shard_once():
    from flashattention_c_b200.ring import ring_schedule

    for world in (1, 2, 4, 8):
        for r in range(world):
            sched = ring_schedule(r, world)
            assert [s for s, _ in sched] == list(range(world))
            assert sched[0][1] == r                     # step 0 is the local shard
            assert sorted(src for _, src in sched) == list(range(world))
        # what rank r holds at step s is what rank r-1 held at step s-1 (send to r+1, receive from r-1)
        for s in range(1, world):
            for r in range(world):
                assert ring_schedule(r, world)[s][1] == ring_schedule((r - 1) % world, world)[s - 1][1]


def test_bench_reference_arm_prints_contract_line_without_gpu():
    """`bench.py --impl reference` must always print one JSON line with the contract keys; on a GPU-less box it times
    the oracle's CPU port on a bounded sample (the CUDA reference arm is exercised on the B200)."""
    import json
    import sys

    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-1000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "impl", "cpu_baseline", "e2e", "config"):
        assert key in line, key
    assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_own_arm_refuses_to_run_without_gpu():
    import sys

    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)


def test_bench_reads_measured_peaks_or_states_the_fallback(tmp_path, monkeypatch):
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) when present — burst for kernels timed alone, sustained
    for the long ring step — else the profiling guide's fallback, labelled as such."""
    import importlib
    import json

    sys.path.insert(0, str(ROOT))
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "ROOT", tmp_path)
    p = bench.load_peaks()
    assert p["src"].startswith("fallback") and p["bf16"] == 1590.0 and p["hbm"] == 6650.0 and p["bf16_sustained"] < p["bf16"]
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"bf16_tflops": 1618.6, "bf16_tflops_sustained": 1367.7, "hbm_gbs": 6555.5}))
    p = bench.load_peaks()
    assert p == {"bf16": 1618.6, "bf16_sustained": 1367.7, "hbm": 6555.5, "src": "MEASURED_PEAKS.json"}
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"bf16_tflops": 1600.0, "hbm_gbs": 6500.0}))
    assert bench.load_peaks()["bf16_sustained"] == 1600.0        # older file without the sustained figure
    assert bench.flops_of(2, 8, 8192, 64) == 4.0 * 16 * 8192 * 8192 * 64 and bench.bytes_of(2, 8, 8192, 64, 4) == 4.0 * 16 * 8192 * 64 * 4


def test_ring_calls_cover_every_visible_pair_once_and_flag_the_last_call():
    """ring_calls: the attention calls of one rank's ring forward.  Every (query chunk, key chunk) pair that is visible is
    covered exactly once, and exactly one call per accumulator — the last — is flagged (it writes the caller's dtype)."""
    from flashattention_c_b200.ring import ring_calls

    for world in (1, 2, 4, 8):
        for rank in range(world):
            calls = ring_calls(rank, world, False, False)
            assert [c[0] for c in calls] == [(rank - s) % world for s in range(world)]
            assert [c[5] for c in calls] == [False] * (world - 1) + [True] and all(c[1] == 0 and c[2] is None for c in calls)
            calls = ring_calls(rank, world, True, False)
            assert sorted(c[0] for c in calls) == list(range(rank + 1)) and calls[0][4] and not any(c[4] for c in calls[1:])
            assert sum(c[5] for c in calls) == 1 and calls[-1][5]
            if world == 1:
                continue
            calls = ring_calls(rank, world, True, True)
            seen = set()
            for src, slot, hq, keys, cz, is_last in calls:
                assert slot == hq
                qc = rank if hq == 0 else 2 * world - 1 - rank
                for kc in ([src] if keys == "lo" else [src, 2 * world - 1 - src]):
                    assert kc <= qc and (cz or kc < qc) and (qc, kc) not in seen
                    seen.add((qc, kc))
            want = {(qc, kc) for qc in (rank, 2 * world - 1 - rank) for kc in range(2 * world) if kc <= qc}
            assert seen == want
            for slot in (0, 1):
                flags = [c[5] for c in calls if c[1] == slot]
                assert flags[-1] and sum(flags) == 1


def test_reference_bench_script_copy_is_pinned_byte_for_byte():
    """tests/golden/bench_flashattention.py.sha256 pins the reference's own caller: the copy the GPU test runs from compat/
    (oracle/_ref/, made by oracle/Makefile) and the reference's file where it is mounted must both hash to it."""
    import hashlib

    want = (ROOT / "tests" / "golden" / "bench_flashattention.py.sha256").read_text().strip()
    seen = 0
    for path in (Path("/root/reference/bench_flashattention.py"), ROOT / "oracle" / "_ref" / "bench_flashattention.py"):
        if path.exists():
            assert hashlib.sha256(path.read_bytes()).hexdigest() == want, path
            seen += 1
    if not seen:
        pytest.skip("neither /root/reference nor oracle/_ref is present on this box")
    # what the script's load(name='flash', sources=['src/main.cpp', 'src/flashattention.cu']) finds when run from compat/
    for f in ("main.cpp", "flashattention.cu"):
        assert (ROOT / "compat" / "src" / f).exists()
    text = (ROOT / "compat" / "src" / "flashattention.cu").read_text()
    assert "__global__" not in text and "dlopen" in text        # a host-only stub over the C-ABI, no kernel of its own


def test_bench_host_helpers_without_a_gpu():
    """bench.py's host-side helpers: core / NUMA binding never raises (it is an optimisation), and the fp64 sampled-row checker
    (torch, used for the `parity` fields of the multi-GPU sections) agrees with the oracle."""
    import importlib
    import os

    import numpy as np

    sys.path.insert(0, str(ROOT))
    bench = importlib.import_module("bench")
    before = os.sched_getaffinity(0)
    try:
        for policy in ("off", "local", "spread"):
            os.environ["FA_BENCH_NUMA"] = policy
            info = bench.pin_to_gpu_numa(0, 2)
            assert info["policy"] == policy
    finally:
        os.environ.pop("FA_BENCH_NUMA", None)
        os.sched_setaffinity(0, before)
    from oracle import fa_oracle

    rng = np.random.default_rng(3)
    q, k, v = (rng.standard_normal(s, dtype=np.float32) for s in ((1, 5, 40, 16), (1, 5, 96, 16), (1, 5, 96, 16)))
    o_ref, lse_ref = fa_oracle.f64(q, k, v, 0.25, False)
    o_bad = o_ref.copy()
    o_bad[0, 3, 7, 2] += 0.5
    tq, tk, tv = (torch.from_numpy(x) for x in (q, k, v))
    err_o, err_l = bench.fp64_rows_check(torch, tq, tk, tv, torch.from_numpy(o_ref), torch.from_numpy(lse_ref), 0.25, 40, 1)
    assert err_o < 1e-12 and err_l < 1e-12
    err_o, _ = bench.fp64_rows_check(torch, tq, tk, tv, torch.from_numpy(o_bad), None, 0.25, 40, 1)
    assert abs(err_o - 0.5) < 1e-9
