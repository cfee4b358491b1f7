"""CPU tests of the oracle itself: internal consistency, the reference's own CPU implementation where it was
compiled (oracle/_ref), and the golden vectors captured from the reference CUDA kernel on a B200."""
import glob
from pathlib import Path

import numpy as np
import pytest

from conftest import seeded

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("bh,nq,nk,d,causal", [
    (2, 64, 64, 64, False), (2, 64, 64, 64, True), (3, 100, 100, 32, False), (3, 100, 100, 32, True),
    (1, 33, 77, 16, False), (2, 40, 72, 8, True), (1, 1, 1, 64, True), (2, 128, 128, 128, True),
])
def test_tiled_matches_f64_and_numpy(oracle, bh, nq, nk, d, causal):
    q, k, v = seeded((bh, nq, d), 1), seeded((bh, nk, d), 2), seeded((bh, nk, d), 3)
    scale = 1.0 / np.sqrt(d)
    o_t, lse_t = oracle.tiled(q, k, v, scale, causal)
    o_f, lse_f = oracle.f64(q, k, v, scale, causal)
    o_n, lse_n = oracle.numpy_f64(q, k, v, scale, causal)
    assert np.abs(o_f - o_n).max() < 1e-12 and np.abs(lse_f - lse_n).max() < 1e-12
    assert np.abs(o_t - o_f).max() < 5e-6
    assert np.abs(lse_t - lse_f).max() < 5e-6


def test_reference_scale_one_semantics(oracle):
    """forward() in the reference passes scaling = 1.0 (src/flashattention.cu:593): peaky softmax, still consistent."""
    q, k, v = seeded((2, 96, 64), 4), seeded((2, 96, 64), 5), seeded((2, 96, 64), 6)
    o_t, _ = oracle.tiled(q, k, v, 1.0, False)
    o_f, _ = oracle.f64(q, k, v, 1.0, False)
    assert np.abs(o_t - o_f).max() < 2e-5


def test_rows_sum_to_one_property(oracle):
    """V = 1 (the reference harness's own input, test.cu:627-631) must give O = 1 everywhere."""
    q, k = seeded((2, 70, 32), 7), seeded((2, 70, 32), 8)
    v = np.ones((2, 70, 32), dtype=np.float32)
    for causal in (False, True):
        o, _ = oracle.tiled(q, k, v, 0.3, causal)
        assert np.abs(o - 1.0).max() < 1e-5


def test_merge_rule_equals_full_attention(oracle):
    """log-sum-exp merge of two disjoint key partitions == attention over all keys (the ring-step rule)."""
    q, k, v = seeded((2, 48, 32), 9), seeded((2, 96, 32), 10), seeded((2, 96, 32), 11)
    o_full, lse_full = oracle.f64(q, k, v, 0.2, False)
    o_a, lse_a = oracle.f64(q, k[:, :40], v[:, :40], 0.2, False)
    o_b, lse_b = oracle.f64(q, k[:, 40:], v[:, 40:], 0.2, False)
    o_m, lse_m = oracle.merge(o_a, lse_a, o_b, lse_b)
    assert np.abs(o_m - o_full).max() < 1e-12 and np.abs(lse_m - lse_full).max() < 1e-12


def test_llmc_cpu_restatement_matches_generic_oracle(oracle):
    B, T, C, NH = 2, 80, 96, 3
    inp = np.random.default_rng(12).random((B, T, 3 * C), dtype=np.float32) * 2 - 1  # make_random_float range, common.h:46-52
    out = oracle.llmc_cpu(inp, B, T, C, NH)
    q, k, v = oracle.packed_qkv_to_bhnd(inp, B, T, C, NH)
    o, _ = oracle.f64(q, k, v, 1.0 / np.sqrt(C // NH), True)
    assert np.abs(out - o.transpose(0, 2, 1, 3).reshape(B, T, C)).max() < 2e-6


def test_pinned_against_reference_cpu_implementation(oracle):
    """The reference's own attention_forward_cpu (src/llm.c/attention_forward.cu:53-125), compiled from
    /root/reference into oracle/_ref/libllmc_ref.so, against our restatement: must agree to the last bit."""
    B, T, C, NH = 2, 64, 128, 4
    inp = np.random.default_rng(13).random((B, T, 3 * C), dtype=np.float32) * 2 - 1
    ref = oracle.ref_llmc_cpu(inp, B, T, C, NH)
    if ref is None:
        pytest.skip("oracle/_ref/libllmc_ref.so not built here (needs /root/reference); pinned by tests/golden instead")
    mine = oracle.llmc_cpu(inp, B, T, C, NH)
    assert np.array_equal(mine, ref)


def _golden_files():
    return sorted(glob.glob(str(GOLDEN / "ref_kernel_*.npz")))


@pytest.mark.parametrize("path", _golden_files() or [None])
def test_pinned_against_reference_kernel_golden(oracle, path):
    """Golden vectors = outputs of the reference CUDA kernel (flash_tiled_coarse[_causal], rebuilt for sm_100a, run on
    a B200 by tests/golden/make_golden.py).  The tile-order restatement must reproduce them to fp32 round-off."""
    if path is None:
        pytest.skip("no golden vectors committed yet")
    g = np.load(path)
    q, k, v, o_ref = g["q"], g["k"], g["v"], g["o"]
    causal = bool(g["causal"])
    o_t, _ = oracle.tiled(q, k, v, 1.0, causal)       # the reference torch path uses scaling = 1.0
    o_f, _ = oracle.f64(q, k, v, 1.0, causal)
    assert np.abs(o_t - o_ref).max() < 2e-5, path
    assert np.abs(o_f - o_ref).max() < 2e-5, path


@pytest.mark.parametrize("causal,hk,nq,nk", [(False, 2, 9, 12), (True, 2, 12, 12), (True, 1, 7, 11), (True, 2, 11, 7)])
def test_backward_oracle_matches_central_differences(oracle, causal, hk, nq, nk):
    """backward_f64 is written out analytically; pin it against central differences of the forward oracle (numpy_f64) contracted
    with dO, on every input element (fp64: agreement to ~1e-7 relative).  Covers grouped K/V heads and n_q != n_k causal, where
    rows without a visible key (n_q > n_k) must get zero gradients."""
    rng = np.random.default_rng(5)
    B, H, d, scale = 1, 2, 4, 0.6
    q = rng.standard_normal((B, H, nq, d))
    k = rng.standard_normal((B, hk, nk, d))
    v = rng.standard_normal((B, hk, nk, d))
    d_o = rng.standard_normal((B, H, nq, d))
    g = H // hk

    def loss(q_, k_, v_):
        o, _ = oracle.numpy_f64(q_, np.repeat(k_, g, axis=1), np.repeat(v_, g, axis=1), scale=scale, causal=causal)
        return float((o * d_o).sum())

    dq, dk, dv = oracle.backward_f64(q, k, v, d_o, scale=scale, causal=causal)
    eps = 1e-6
    for name, x, gx in (("q", q, dq), ("k", k, dk), ("v", v, dv)):
        num = np.zeros_like(x)
        it = np.nditer(x, flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            old = x[idx]
            x[idx] = old + eps
            up = loss(q, k, v)
            x[idx] = old - eps
            dn = loss(q, k, v)
            x[idx] = old
            num[idx] = (up - dn) / (2 * eps)
        assert np.abs(num - gx).max() < 1e-6 * max(1.0, np.abs(gx).max()), name


@pytest.mark.parametrize("causal,nq,nk", [(False, 40, 56), (True, 48, 48), (True, 30, 50), (True, 50, 30)])
def test_blockwise_recomputation_backward_matches_the_fp64_oracle(oracle, causal, nq, nk):
    """`autograd.recompute_backward` (plain torch matmuls over row blocks: the gradient path of fp32 tensors and the checker of the
    backward kernels at lengths no N x N reference fits) against backward_f64, on CPU tensors from the forward oracle's O and LSE."""
    import torch

    from flashattention_c_b200.autograd import recompute_backward

    rng = np.random.default_rng(11)
    B, H, d, scale = 2, 3, 16, 0.25
    q, do = (rng.standard_normal((B, H, nq, d)).astype(np.float32) for _ in range(2))
    k, v = (rng.standard_normal((B, H, nk, d)).astype(np.float32) for _ in range(2))
    o, lse = oracle.numpy_f64(q, k, v, scale=scale, causal=causal)
    got = recompute_backward(*(torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)) for t in (q, k, v, o)),
                             torch.from_numpy(lse.astype(np.float32)), torch.from_numpy(do), causal, scale)
    want = oracle.backward_f64(q, k, v, do, scale=scale, causal=causal)
    for name, g, w in zip(("dq", "dk", "dv"), got, want):
        assert np.abs(g.numpy() - w).max() < 2e-5 * max(1.0, np.abs(w).max()), name
