// compat/src/main.cpp — the pybind half of the reference's torch extension, for running the reference's own
// bench_flashattention.py UNMODIFIED against the B200 kernels:
//
//     cd compat && python /path/to/FlashAttention.C/bench_flashattention.py [--batch_size B --seq_len N --masking 1]
//
// That script JIT-builds `load(name='flash', sources=['src/main.cpp', 'src/flashattention.cu'], extra_cuda_cflags=['-O3'])`
// relative to the working directory (bench_flashattention.py:10) and calls `.forward(q, k, v, masking)` (line 70); this file
// and src/flashattention.cu next to it are what it finds here.  Stands in for src/main.cpp:1-6 of the reference: the same
// module name and the same exported function, nothing else.
#include <torch/extension.h>

// defined in src/flashattention.cu (this directory): validates, then calls libfa_b200.so through its C-ABI
torch::Tensor forward(torch::Tensor Q_d, torch::Tensor K_d, torch::Tensor V_d, bool causal);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "FlashAttention.C operator surface over libfa_b200.so (sm_100a: TMA + tcgen05 + TMEM)";
  m.def("forward", &forward, "forward(Q, K, V, causal) -> O, scores not scaled (reference semantics)", pybind11::arg("Q_d"),
        pybind11::arg("K_d"), pybind11::arg("V_d"), pybind11::arg("causal"));
}
