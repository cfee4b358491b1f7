"""GPU parity tests (-m gpu) of the backward kernels (include/fa_b200.h: fa_backward; csrc/fa_bwd_sm100.cuh), called through the
C-ABI, against the fp64 oracle (oracle.backward_f64, itself pinned by central differences in tests/test_oracle.py) at sizes it
finishes in seconds, and against a float64 torch.autograd evaluation on the GPU at the long ones.

Tolerance.  The kernels see the 16-bit inputs exactly, accumulate in fp32, and round P and dS = P (dP - D) to the 16-bit
operand type before the dV / dK / dQ contractions (as every 16-bit FlashAttention backward does), so a gradient element carries
the half-ulp of those operands averaged over the contraction: bf16 2^-9 -> measured 2e-3 .. 6e-3 of the tensor's largest
magnitude (profiles/r02_bwd_bringup_check.log), fp16 2^-12 -> 3e-4 .. 6e-4.  Gates: 2e-2 (bf16), 4e-3 (fp16) of max(1, max |ref|).
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = {torch.bfloat16: 2e-2, torch.float16: 4e-3}


def _inputs(B, H, Hk, nq, nk, d, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, nq, d, generator=g).to(dtype)
    k = torch.randn(B, Hk, nk, d, generator=g).to(dtype)
    v = torch.randn(B, Hk, nk, d, generator=g).to(dtype)
    do = torch.randn(B, H, nq, d, generator=g).to(dtype)
    return q, k, v, do


def _kernel(fab, dev, q, k, v, do, causal, scale, fwd=None):
    qd, kd, vd, dod = (t.to(dev) for t in (q, k, v, do))
    o, lse = fwd if fwd is not None else fab.attention(qd, kd, vd, causal=causal, scale=scale, return_lse=True)
    before = fab.launch_count()
    dq, dk, dv = fab.attention_backward(qd, kd, vd, o, lse, dod, causal=causal, scale=scale)
    torch.cuda.synchronize()
    assert fab.launch_count() == before + 3          # statistics + dK/dV + dQ: the native kernels are what ran
    assert fab.last_impl() == fab.FA_IMPL_TCGEN05
    return dq, dk, dv


def _rel(got, ref):
    got = got.double().cpu().numpy() if isinstance(got, torch.Tensor) else got
    ref = ref.double().cpu().numpy() if isinstance(ref, torch.Tensor) else ref
    assert np.isfinite(got).all()
    # relative to the tensor's largest magnitude, or to 1 where the exact gradient is (nearly) zero — e.g. a single key: P = 1 and
    # dP - D = 0 in exact arithmetic, while the kernel's D comes from the 16-bit O the forward stored
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1.0))


def _torch_f64(q, k, v, do, causal, scale, dev):
    """float64 autograd on the GPU (grouped K/V heads expanded with repeat_interleave, so their gradients sum over the group)."""
    q64, k64, v64 = (t.to(dev).double().requires_grad_(True) for t in (q, k, v))
    g = q.shape[1] // k.shape[1]
    kk, vv = k64.repeat_interleave(g, dim=1), v64.repeat_interleave(g, dim=1)
    nq, nk = q.shape[2], k.shape[2]
    s = (q64 @ kk.transpose(-1, -2)) * scale
    if causal:
        i, j = torch.arange(nq, device=dev)[:, None], torch.arange(nk, device=dev)[None, :]
        s = s.masked_fill(j > i + (nk - nq), float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)     # rows without a visible key: O = 0, no gradient
    (p @ vv).backward(do.to(dev).double())
    return q64.grad, k64.grad, v64.grad


CASES = [
    # B, H, Hk, nq, nk, d, causal, dtype
    (1, 1, 1, 128, 128, 64, False, torch.bfloat16),      # one resident tile, two streamed tiles
    (1, 1, 1, 128, 128, 128, False, torch.bfloat16),
    (1, 2, 2, 256, 256, 64, True, torch.bfloat16),
    (1, 2, 2, 200, 333, 64, True, torch.bfloat16),       # ragged both ways, causal offset > 0
    (2, 4, 2, 192, 320, 128, False, torch.bfloat16),     # grouped K/V heads: dK, dV summed over the group inside the CTA
    (1, 2, 1, 77, 130, 128, True, torch.float16),        # multi-query, fp16
    (1, 3, 3, 300, 100, 64, True, torch.bfloat16),       # n_q > n_k causal: the first 200 rows see no key (zero gradients)
    (1, 2, 2, 640, 640, 96, True, torch.bfloat16),       # head dim below the instance: zero-padded by TMA, clipped on store
    (1, 2, 2, 1024, 1024, 32, False, torch.float16),
    (1, 2, 2, 1100, 900, 128, False, torch.bfloat16),    # more streamed steps than ring stages
    (1, 1, 1, 1, 1, 8, True, torch.bfloat16),            # smallest problem
    (1, 1, 1, 1, 700, 64, False, torch.float16),         # decode-like: one query row
    (1, 1, 1, 700, 1, 64, False, torch.bfloat16),        # one key
]


@pytest.mark.parametrize("B,H,Hk,nq,nk,d,causal,dtype", CASES)
def test_backward_vs_fp64_oracle(fab, oracle, cuda_device, B, H, Hk, nq, nk, d, causal, dtype):
    q, k, v, do = _inputs(B, H, Hk, nq, nk, d, dtype, seed=nq * 7 + nk)
    scale = 1.0 / math.sqrt(d)
    dq, dk, dv = _kernel(fab, cuda_device, q, k, v, do, causal, scale)
    rq, rk, rv = oracle.backward_f64(q.float().numpy(), k.float().numpy(), v.float().numpy(), do.float().numpy(), scale=scale, causal=causal)
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert got.dtype == dtype
        assert _rel(got, ref) < TOL[dtype], name


def test_rows_without_a_visible_key_get_exactly_zero_gradient(fab, cuda_device):
    """Causal with n_q > n_k: rows 0 .. n_q - n_k - 1 see no key (forward: O = 0, LSE = -inf); their dQ is exactly 0 and they add
    nothing to dK, dV — including 128-row tiles in which NO row sees a key (no streamed step at all: the epilogue writes zeros
    without touching TMEM)."""
    q, k, v, do = _inputs(1, 2, 2, 700, 300, 64, torch.bfloat16, seed=3)
    dq, dk, dv = _kernel(fab, cuda_device, q, k, v, do, True, 0.125)
    assert float(dq[:, :, :400].abs().max()) == 0.0
    rq, rk, rv = _torch_f64(q, k, v, do, True, 0.125, cuda_device)
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        assert _rel(got, ref) < TOL[torch.bfloat16]


@pytest.mark.parametrize("d,dtype,causal,n", [(128, torch.bfloat16, False, 4096), (128, torch.bfloat16, True, 4096), (64, torch.float16, True, 4096),
                                              (64, torch.bfloat16, False, 3000)])
def test_backward_long_sequences_vs_float64_autograd(fab, cuda_device, d, dtype, causal, n):
    """Tens of streamed steps per CTA (the TMA ring and both TMEM buffers wrap many times), every CTA of a multi-wave grid."""
    q, k, v, do = _inputs(2, 4, 4, n, n, d, dtype, seed=n + d)
    scale = 1.0 / math.sqrt(d)
    dq, dk, dv = _kernel(fab, cuda_device, q, k, v, do, causal, scale)
    rq, rk, rv = _torch_f64(q, k, v, do, causal, scale, cuda_device)
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert _rel(got, ref) < TOL[dtype], name


def test_backward_is_bit_reproducible(fab, cuda_device):
    """No atomics anywhere: two runs give identical bits, and a (batch, head) slice computed alone equals the same slice inside the batch."""
    q, k, v, do = _inputs(2, 4, 2, 900, 1300, 128, torch.bfloat16, seed=11)
    o, lse = fab.attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device), causal=True, scale=0.09, return_lse=True)
    a = _kernel(fab, cuda_device, q, k, v, do, True, 0.09, fwd=(o, lse))
    b = _kernel(fab, cuda_device, q, k, v, do, True, 0.09, fwd=(o, lse))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    sl = _kernel(fab, cuda_device, q[1:, 2:4], k[1:, 1:2], v[1:, 1:2], do[1:, 2:4], True, 0.09, fwd=(o[1:, 2:4], lse[1:, 2:4]))
    assert torch.equal(sl[0], a[0][1:, 2:4]) and torch.equal(sl[1], a[1][1:, 1:2]) and torch.equal(sl[2], a[2][1:, 1:2])


def test_backward_on_strided_views_of_a_packed_qkv_tensor(fab, cuda_device):
    """Q, K, V as head-interleaved views of one (B, T, 3, NH, hs) tensor (the llm.c layout, src/llm.c/attention_forward.cu:1106-1179)
    and a dO that is a transposed view: the strides go into the tensor maps, nothing is copied."""
    B, T, NH, hs = 2, 384, 3, 64
    g = torch.Generator().manual_seed(5)
    packed = torch.randn(B, T, 3, NH, hs, generator=g).to(torch.bfloat16).to(cuda_device)
    q, k, v = (packed[:, :, i].permute(0, 2, 1, 3) for i in range(3))          # [B, NH, T, hs] views
    do = torch.randn(B, T, NH, hs, generator=g).to(torch.bfloat16).to(cuda_device).permute(0, 2, 1, 3)
    assert not q.is_contiguous() and not do.is_contiguous()
    o, lse = fab.attention(q, k, v, causal=True, scale=0.125, return_lse=True)
    dq, dk, dv = fab.attention_backward(q, k, v, o, lse, do, causal=True, scale=0.125)
    rq, rk, rv = _torch_f64(q.cpu(), k.cpu(), v.cpu(), do.cpu(), True, 0.125, cuda_device)
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        assert _rel(got, ref) < TOL[torch.bfloat16]


def test_autograd_uses_the_backward_kernels(fab, cuda_device):
    """attention_autograd on bf16 tensors: forward kernel + the three backward launches (no torch matmul path), 3-D [B*H, N, d] form,
    gradients vs float64 autograd."""
    g = torch.Generator().manual_seed(9)
    q, k, v = (torch.randn(6, n, 64, generator=g).to(torch.bfloat16).to(cuda_device).requires_grad_(True) for n in (520, 520, 520))
    d_o = torch.randn(6, 520, 64, generator=g).to(torch.bfloat16).to(cuda_device)
    o = fab.attention_autograd(q, k, v, causal=True)
    before = fab.launch_count()
    o.backward(d_o)
    torch.cuda.synchronize()
    assert fab.launch_count() == before + 3
    rq, rk, rv = _torch_f64(q.detach()[None].cpu(), k.detach()[None].cpu(), v.detach()[None].cpu(), d_o[None].cpu(), True, 0.125, cuda_device)
    for got, ref in ((q.grad, rq[0]), (k.grad, rk[0]), (v.grad, rv[0])):
        assert _rel(got, ref) < TOL[torch.bfloat16]


def test_backward_unsupported_cases_raise(fab, cuda_device):
    q = torch.randn(2, 64, 64, device=cuda_device)
    o, lse = fab.attention(q, q, q, return_lse=True)
    with pytest.raises(fab.FaError):
        fab.attention_backward(q, q, q, o, lse, o)            # fp32: no instance (attention_autograd takes the recomputation path)
    h = torch.randn(2, 64, 256, device=cuda_device).to(torch.bfloat16)
    oh, lh = fab.attention(h, h, h, return_lse=True)
    with pytest.raises(fab.FaError):
        fab.attention_backward(h, h, h, oh, lh, oh)           # head dim 256: two accumulators of 256 columns exceed TMEM
    b = torch.randn(2, 64, 64, device=cuda_device).to(torch.bfloat16)
    ob, lb = fab.attention(b, b, b, return_lse=True)
    with pytest.raises(fab.FaError):
        fab.attention_backward(b, b, b, ob, lb, ob[:, :32])   # dO shape


@pytest.mark.parametrize("seed", [1, 2])
def test_backward_randomised_shapes(fab, cuda_device, seed):
    """2 x 40 seeded random problems (batch, heads, K/V heads, n_q, n_k, head dim, dtype, causal) vs float64 autograd on the GPU."""
    rng = np.random.default_rng(seed)
    worst = 0.0
    for case in range(40):
        dtype = [torch.bfloat16, torch.float16][int(rng.integers(0, 2))]
        d = int(rng.choice([8, 16, 32, 40, 64, 72, 96, 128]))
        Hk = int(rng.integers(1, 4))
        H = Hk * int(rng.integers(1, 4))
        B = int(rng.integers(1, 3))
        nq = int(rng.integers(1, 700))
        nk = nq if rng.random() < 0.5 else int(rng.integers(1, 900))
        causal = bool(rng.integers(0, 2))
        q, k, v, do = _inputs(B, H, Hk, nq, nk, d, dtype, seed=1000 * seed + case)
        scale = float(rng.choice([1.0 / math.sqrt(d), 0.3]))
        dq, dk, dv = _kernel(fab, cuda_device, q, k, v, do, causal, scale)
        rq, rk, rv = _torch_f64(q, k, v, do, causal, scale, cuda_device)
        for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
            e = _rel(got, ref)
            worst = max(worst, e / TOL[dtype])
            assert e < TOL[dtype], (case, name, (B, H, Hk, nq, nk, d, dtype, causal, scale), e)
    print(f"worst error / tolerance over 40 cases: {worst:.3f}")


def test_forward_and_backward_under_cuda_graph_capture(fab, cuda_device):
    """A training step's attention — forward with LSE, then the three backward launches — captured once and replayed on new data.
    Under capture the backward's statistics workspace is a stream-ordered allocation of the graph (cudaMallocAsync / cudaFreeAsync
    nodes) instead of the per-stream cached buffer."""
    g0 = torch.Generator().manual_seed(21)
    q, k, v, do = (torch.randn(2, 4, 600, 64, generator=g0).to(torch.bfloat16).to(cuda_device) for _ in range(4))
    o, lse = fab.attention(q, k, v, causal=True, return_lse=True)                 # warm-up outside capture
    fab.attention_backward(q, k, v, o, lse, do, causal=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            o_c, lse_c = fab.attention(q, k, v, causal=True, return_lse=True)
            grads_c = fab.attention_backward(q, k, v, o_c, lse_c, do, causal=True)
    torch.cuda.current_stream().wait_stream(s)
    for t in (q, k, v, do):                                                       # new data in the captured buffers
        t.copy_(torch.randn(t.shape, generator=g0).to(torch.bfloat16))
    graph.replay()
    torch.cuda.synchronize()
    o_e, lse_e = fab.attention(q, k, v, causal=True, return_lse=True)
    grads_e = fab.attention_backward(q, k, v, o_e, lse_e, do, causal=True)
    torch.cuda.synchronize()
    assert torch.equal(o_c, o_e)
    for a, b in zip(grads_c, grads_e):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,causal", [(32768, True), (16384, False)])
def test_backward_at_ring_scale_lengths_vs_blockwise_recomputation(fab, cuda_device, n, causal):
    """Hundreds of streamed steps per CTA (N = 32768: 256) — lengths at which no N x N reference fits.  Checker: the blockwise
    recomputation backward (torch matmuls in fp32 over row blocks; itself checked against the fp64 oracle on CPU tensors by
    tests/test_oracle.py::test_blockwise_recomputation_backward_matches_the_fp64_oracle), on the same O and LSE."""
    from flashattention_c_b200.autograd import recompute_backward

    q, k, v, do = (t.to(cuda_device) for t in _inputs(1, 2, 2, n, n, 128, torch.bfloat16, seed=n))
    scale = 1.0 / math.sqrt(128)
    o, lse = fab.attention(q, k, v, causal=causal, scale=scale, return_lse=True)
    got = fab.attention_backward(q, k, v, o, lse, do, causal=causal, scale=scale)
    want = recompute_backward(q, k, v, o, lse, do, causal, scale)
    for name, a, b in zip(("dq", "dk", "dv"), got, want):
        assert _rel(a, b) < TOL[torch.bfloat16], name


def test_backward_under_compute_sanitizer_timing_stretch():
    """The bring-up cases (incl. grouped K/V heads with more streamed steps than ring slots) under compute-sanitizer memcheck: no
    memory error, and — because the tool stretches every window between a hand-over and its consumer — still the right
    numbers.  This is the run that exposed a ring slot being released one contraction too early (its per-row statistics were read
    by the element-wise warps after the tile's last reader on the tensor pipe had been issued): correct at full speed, wrong dK
    under the sanitizer."""
    import shutil
    import subprocess
    import sys
    from pathlib import Path

    tool = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not Path(tool).exists():
        pytest.skip("compute-sanitizer not installed")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([tool, "--tool", "memcheck", "--error-exitcode", "9", sys.executable, str(root / "scripts" / "bwd_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BAD 0" in r.stdout and "ERROR SUMMARY: 0 errors" in r.stdout
