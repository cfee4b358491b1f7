// compat/src/flashattention.cu — what bench_flashattention.py:10 compiles as 'src/flashattention.cu' when it is run from
// compat/: the reference's `forward(Q_d, K_d, V_d, causal)` (src/flashattention.cu:603-617 of the reference) as a thin host
// stub over the C-ABI of libfa_b200.so.  There is no device code in this file — the kernels live in the prebuilt library,
// which is located relative to this source file and opened with dlopen (the script passes no linker flags, so nothing can
// be linked at build time).  Semantics kept from the reference: scores are not scaled (scaling = 1.0, lines 593/600), the
// result is a fresh [B*H, N, d] tensor on the inputs' device.  Unlike the reference, inputs are validated (it only asserts
// size(2) == 64), 4-D [B, H, N, d] is accepted too, the launch goes to the current torch stream, and errors are raised.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <dlfcn.h>
#include <torch/types.h>

#include <cstdlib>
#include <mutex>
#include <string>

#include "../../include/fa_b200.h"

namespace {

struct Abi {
  decltype(&fa_forward) forward = nullptr;
  decltype(&fa_strerror) strerror_ = nullptr;
  decltype(&fa_last_cuda_error) last_cuda_error = nullptr;
  std::string error;
};

// <repo>/compat/src/flashattention.cu -> <repo>/flashattention.c_b200/libfa_b200.so ($FA_B200_LIB overrides)
std::string library_path() {
  if (const char* e = getenv("FA_B200_LIB")) return e;
  std::string here = __FILE__;
  for (int up = 0; up < 3; ++up) {
    const size_t cut = here.find_last_of('/');
    here = cut == std::string::npos ? std::string(".") : here.substr(0, cut);
  }
  return here + "/flashattention.c_b200/libfa_b200.so";
}

const Abi& abi() {
  static Abi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const std::string path = library_path();
    void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      a.error = std::string("cannot open ") + path + ": " + dlerror() + " (build it with `python flashattention.c_b200/build.py`; there is no fallback)";
      return;
    }
    a.forward = reinterpret_cast<decltype(a.forward)>(dlsym(h, "fa_forward"));
    a.strerror_ = reinterpret_cast<decltype(a.strerror_)>(dlsym(h, "fa_strerror"));
    a.last_cuda_error = reinterpret_cast<decltype(a.last_cuda_error)>(dlsym(h, "fa_last_cuda_error"));
    if (!a.forward || !a.strerror_ || !a.last_cuda_error) a.error = path + " does not export the fa_b200 C-ABI";
  });
  return a;
}

}  // namespace

torch::Tensor forward(torch::Tensor Q_d, torch::Tensor K_d, torch::Tensor V_d, bool causal) {
  const Abi& a = abi();
  TORCH_CHECK(a.error.empty(), "flash.forward: ", a.error);
  TORCH_CHECK(Q_d.is_cuda() && K_d.is_cuda() && V_d.is_cuda(), "flash.forward: Q, K, V must be CUDA tensors (there is no CPU path)");
  TORCH_CHECK(Q_d.device() == K_d.device() && Q_d.device() == V_d.device(), "flash.forward: Q, K, V must be on one device");
  TORCH_CHECK(Q_d.scalar_type() == K_d.scalar_type() && Q_d.scalar_type() == V_d.scalar_type(), "flash.forward: dtype mismatch");
  const auto st = Q_d.scalar_type();
  TORCH_CHECK(st == torch::kFloat32 || st == torch::kBFloat16 || st == torch::kFloat16, "flash.forward: float32, bfloat16 or float16 only");
  TORCH_CHECK((Q_d.dim() == 3 || Q_d.dim() == 4) && K_d.dim() == Q_d.dim() && V_d.dim() == Q_d.dim(),
              "flash.forward: expected [B*H, N, d] or [B, H, N, d]");
  Q_d = Q_d.contiguous();
  K_d = K_d.contiguous();
  V_d = V_d.contiguous();
  const int r = Q_d.dim();
  const int64_t batch = r == 4 ? Q_d.size(0) : 1, heads = Q_d.size(r - 3), n_q = Q_d.size(r - 2), d = Q_d.size(r - 1);
  const int64_t n_k = K_d.size(r - 2);
  TORCH_CHECK(K_d.size(r - 3) == heads && V_d.size(r - 3) == heads && (r == 3 || (K_d.size(0) == batch && V_d.size(0) == batch)) &&
                  V_d.size(r - 2) == n_k && K_d.size(r - 1) == d && V_d.size(r - 1) == d,
              "flash.forward: Q / K / V shape mismatch");
  c10::cuda::CUDAGuard guard(Q_d.device());
  torch::Tensor O = torch::empty_like(Q_d);
  const int dtype = st == torch::kBFloat16 ? FA_BF16 : (st == torch::kFloat16 ? FA_F16 : FA_F32);
  const int rc = a.forward(Q_d.data_ptr(), K_d.data_ptr(), V_d.data_ptr(), O.data_ptr(), nullptr, batch, heads, n_q, n_k, (int32_t)d,
                           /*scale=*/1.0f, causal ? 1 : 0, dtype, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == FA_OK, "flash.forward failed: ", a.strerror_(rc), " ", a.last_cuda_error());
  return O;
}
