// torch_binding.cpp — the torch-extension surface of the reference, rebuilt over the C-ABI.
//
// Drop-in for src/main.cpp:3-6 + forward() in src/flashattention.cu:603-617:
//     module.forward(Q, K, V, causal) -> O
// with the reference's semantics (scale = 1.0, src/flashattention.cu:593/600; 3-D [B*H, N, d] tensors) and
// without its hazards: inputs are validated (the reference only asserts size(2) == 64), 4-D [B, H, N, d] is
// accepted, the launch goes to the current torch stream and nothing synchronises the device.
// Compiled with g++ only (no device code here); links libfa_b200.so.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cmath>
#include <tuple>

#include "../../include/fa_b200.h"

namespace {

struct Shape { int64_t b, h, nq, nk, d; };

Shape check_inputs(const torch::Tensor& Q, const torch::Tensor& K, const torch::Tensor& V) {
  TORCH_CHECK(Q.is_cuda() && K.is_cuda() && V.is_cuda(), "fa_b200.forward: Q, K, V must be CUDA tensors (there is no CPU path)");
  TORCH_CHECK(Q.device() == K.device() && Q.device() == V.device(), "fa_b200.forward: Q, K, V must be on the same device");
  TORCH_CHECK(Q.scalar_type() == K.scalar_type() && Q.scalar_type() == V.scalar_type(), "fa_b200.forward: dtype mismatch");
  TORCH_CHECK(Q.scalar_type() == torch::kFloat32 || Q.scalar_type() == torch::kBFloat16 || Q.scalar_type() == torch::kFloat16,
              "fa_b200.forward: only float32 (tf32 tensor cores), bfloat16 and float16 are supported");
  TORCH_CHECK(Q.dim() == 3 || Q.dim() == 4, "fa_b200.forward: expected [B*H, N, d] or [B, H, N, d]");
  TORCH_CHECK(K.dim() == Q.dim() && V.dim() == Q.dim(), "fa_b200.forward: rank mismatch");
  TORCH_CHECK(Q.is_contiguous() && K.is_contiguous() && V.is_contiguous(), "fa_b200.forward: tensors must be contiguous");
  Shape s;
  if (Q.dim() == 3) {
    s.b = 1; s.h = Q.size(0); s.nq = Q.size(1); s.d = Q.size(2);
    TORCH_CHECK(K.size(0) == s.h && V.size(0) == s.h, "fa_b200.forward: batch*heads mismatch");
    s.nk = K.size(1);
    TORCH_CHECK(V.size(1) == s.nk && K.size(2) == s.d && V.size(2) == s.d, "fa_b200.forward: K/V shape mismatch");
  } else {
    s.b = Q.size(0); s.h = Q.size(1); s.nq = Q.size(2); s.d = Q.size(3);
    TORCH_CHECK(K.size(0) == s.b && V.size(0) == s.b && K.size(1) == s.h && V.size(1) == s.h, "fa_b200.forward: batch/heads mismatch");
    s.nk = K.size(2);
    TORCH_CHECK(V.size(2) == s.nk && K.size(3) == s.d && V.size(3) == s.d, "fa_b200.forward: K/V shape mismatch");
  }
  return s;
}

std::tuple<torch::Tensor, torch::Tensor> forward_impl(const torch::Tensor& Q, const torch::Tensor& K, const torch::Tensor& V,
                                                      bool causal, double scale, bool want_lse) {
  const Shape s = check_inputs(Q, K, V);
  c10::cuda::CUDAGuard guard(Q.device());
  torch::Tensor O = torch::empty_like(Q);
  torch::Tensor lse;
  if (want_lse) {
    auto opts = Q.options().dtype(torch::kFloat32);
    lse = Q.dim() == 3 ? torch::empty({s.h, s.nq}, opts) : torch::empty({s.b, s.h, s.nq}, opts);
  }
  const int dtype = Q.scalar_type() == torch::kBFloat16 ? FA_BF16 : (Q.scalar_type() == torch::kFloat16 ? FA_F16 : FA_F32);
  cudaStream_t st = at::cuda::getCurrentCUDAStream();
  const int rc = fa_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), want_lse ? lse.data_ptr<float>() : nullptr, s.b,
                            s.h, s.nq, s.nk, (int)s.d, (float)scale, causal ? 1 : 0, dtype, st);
  TORCH_CHECK(rc == FA_OK, "fa_b200.forward failed: ", fa_strerror(rc), " ", fa_last_cuda_error());
  return {O, lse};
}

}  // namespace

// same signature as the reference declaration (src/main.cpp:3)
torch::Tensor forward(torch::Tensor Q_d, torch::Tensor K_d, torch::Tensor V_d, bool causal) {
  return std::get<0>(forward_impl(Q_d, K_d, V_d, causal, 1.0, false));
}

std::tuple<torch::Tensor, torch::Tensor> forward_ex(torch::Tensor Q, torch::Tensor K, torch::Tensor V, bool causal, double scale,
                                                    bool return_lse) {
  return forward_impl(Q, K, V, causal, scale, return_lse);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200-native FlashAttention forward (tcgen05/TMA) behind the FlashAttention.C operator surface";
  m.def("forward", &forward, "forward(Q, K, V, causal) -> O   [reference semantics: scale 1.0]", pybind11::arg("Q"),
        pybind11::arg("K"), pybind11::arg("V"), pybind11::arg("causal") = false);
  m.def("forward_ex", &forward_ex, "forward_ex(Q, K, V, causal, scale, return_lse) -> (O, LSE or None)", pybind11::arg("Q"),
        pybind11::arg("K"), pybind11::arg("V"), pybind11::arg("causal") = false, pybind11::arg("scale") = 1.0,
        pybind11::arg("return_lse") = false);
}
