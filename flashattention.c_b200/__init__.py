"""flashattention.c_b200 — B200-native FlashAttention forward behind the FlashAttention.C operator surface.

Host-side mirror of the reference interface (names, argument meaning, error behaviour):

  forward(Q, K, V, causal=False) -> O            src/main.cpp:3-6 / src/flashattention.cu:603-617 (scale 1.0)
  attention(Q, K, V, causal, scale, ...)         the same operator with an explicit scale (default 1/sqrt(d)) and LSE
  attention_forward(kernel_num, out, inp, ...)   src/llm.c/attention_forward.cu:1183-1211 (packed QKV, causal, 1/sqrt(hs))
  load_extension()                               the pybind module (`.forward(Q,K,V,causal)`) bench_flashattention.py:10,70 expects
  attention_backward(Q, K, V, O, LSE, dO, ...)   dQ, dK, dV through the tcgen05 backward kernels (bf16 / fp16, d <= 128; the reference is
                                                 forward only: SURVEY 8 (f4))
  attention_autograd(Q, K, V, causal, scale)     the forward kernel with gradients (backward: the kernels above, or blockwise
                                                 recomputation with torch matmuls for fp32 / d = 256)

Everything runs on hand-written sm_100a kernels through the C-ABI library libfa_b200.so
(include/fa_b200.h).  There is no CPU or eager-PyTorch fallback: without the library or a B200 the calls raise.
"""
from ._lib import FA_BF16, FA_F16, FA_F32, FA_FLAG_BATCH_INVARIANT, FA_IMPL_AUTO, FA_IMPL_SIMT, FA_IMPL_TCGEN05, FaError, lib  # noqa: F401
from .api import (  # noqa: F401
    attention,
    attention_backward,
    attention_forward,
    attention_forward6,
    attention_host,
    forward,
    last_impl,
    launch_count,
    load_extension,
    merge_partials,
)
from .autograd import attention_autograd  # noqa: F401
from .ring import bh_shard_range, ring_attention, ring_p2p_release, sharded_attention, zigzag_shard, zigzag_step_plan, zigzag_unshard  # noqa: F401

__version__ = "0.1.0"
