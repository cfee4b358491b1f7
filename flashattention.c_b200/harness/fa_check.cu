// fa_check.cu — torch-free correctness + timing harness over the C-ABI (include/fa_b200.h).
//   ./fa_check <f32|bf16|f16> <d> <B*H> <N> <causal 0|1> <scale (0 = 1/sqrt(d))> [reps=20] [check_simt=1]
// For each run it (1) launches the tcgen05 path and the SIMT path on the same seeded N(0,1) inputs and
// reports their max-abs difference over the whole tensor, (2) checks sampled rows of both against an fp64
// host evaluation of softmax(scale*q.K^T [+causal]) V, (3) times the tcgen05 path with CUDA events
// (L2 flushed between repetitions; FA_CHECK_TIME_IMPL=2 times the CUDA-core kernel instead) and prints one JSON line.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/fa_b200.h"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      uint32_t wd__[4] = {0, 0, 0, 0};                                               \
      fa_watchdog_info(wd__);                                                        \
      printf("CUDA error %s at %s:%d (watchdog tag %u block %u thread %u parity %u)\n", cudaGetErrorString(e), __FILE__, __LINE__, \
             wd__[0], wd__[1], wd__[2], wd__[3]);                                   \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static inline double urand() {
  g_state = g_state * 6364136223846793005ull + 1442695040888963407ull;
  return ((g_state >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline float nrand() { return (float)(sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand())); }

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float f16_round(float x) { return __half2float(__float2half_rn(x)); }

int main(int argc, char** argv) {
  if (argc < 7) {
    printf("usage: %s f32|bf16|f16 d BH N causal scale [reps] [check_simt]\n", argv[0]);
    return 1;
  }
  const bool f16 = !strcmp(argv[1], "f16");
  const bool bf16 = !strcmp(argv[1], "bf16") || f16;   // "16-bit operands" from here on; f16 picks the conversions
  const int d = atoi(argv[2]);
  const int64_t BH = atoll(argv[3]), N = atoll(argv[4]);
  const int causal = atoi(argv[5]);
  float scale = (float)atof(argv[6]);
  if (scale == 0.f) scale = 1.0f / sqrtf((float)d);
  const int reps = argc > 7 ? atoi(argv[7]) : 20;
  const int check_simt = argc > 8 ? atoi(argv[8]) : 1;
  const size_t n_el = (size_t)BH * N * d, es = bf16 ? 2 : 4;

  std::vector<float> hq(n_el), hk(n_el), hv(n_el);
  for (auto& x : hq) x = nrand();
  for (auto& x : hk) x = nrand();
  for (auto& x : hv) x = nrand();
  if (bf16) {
    for (auto& x : hq) x = f16 ? f16_round(x) : bf16_round(x);
    for (auto& x : hk) x = f16 ? f16_round(x) : bf16_round(x);
    for (auto& x : hv) x = f16 ? f16_round(x) : bf16_round(x);
  }
  auto upload = [&](const std::vector<float>& h, void** dptr) -> int {
    CK(cudaMalloc(dptr, n_el * es));
    if (!bf16) {
      CK(cudaMemcpy(*dptr, h.data(), n_el * 4, cudaMemcpyHostToDevice));
    } else {
      std::vector<uint16_t> t(n_el);
      for (size_t i = 0; i < n_el; ++i) {
        if (f16) { const __half x = __float2half_rn(h[i]); memcpy(&t[i], &x, 2); }
        else { const __nv_bfloat16 x = __float2bfloat16_rn(h[i]); memcpy(&t[i], &x, 2); }
      }
      CK(cudaMemcpy(*dptr, t.data(), n_el * 2, cudaMemcpyHostToDevice));
    }
    return 0;
  };
  void *dq, *dk, *dv, *do_tc, *do_simt;
  float *dlse_tc, *dlse_simt;
  if (upload(hq, &dq) || upload(hk, &dk) || upload(hv, &dv)) return 2;
  CK(cudaMalloc(&do_tc, n_el * es));
  CK(cudaMalloc(&do_simt, n_el * es));
  CK(cudaMalloc(&dlse_tc, BH * N * 4));
  CK(cudaMalloc(&dlse_simt, BH * N * 4));
  CK(cudaMemset(do_tc, 0xff, n_el * es));
  CK(cudaMemset(do_simt, 0xff, n_el * es));

  fa_params p;
  memset(&p, 0, sizeof(p));
  p.q = dq; p.k = dk; p.v = dv; p.batch = 1; p.heads = BH; p.n_q = N; p.n_k = N; p.head_dim = d;
  p.dtype = f16 ? FA_F16 : (bf16 ? FA_BF16 : FA_F32); p.causal = causal; p.scale = scale;
  p.q_stride_n = p.k_stride_n = p.v_stride_n = p.o_stride_n = d;
  p.q_stride_h = p.k_stride_h = p.v_stride_h = p.o_stride_h = N * d;
  p.q_stride_b = p.k_stride_b = p.v_stride_b = p.o_stride_b = BH * N * d;

  p.o = do_tc; p.lse = dlse_tc; p.impl = FA_IMPL_TCGEN05;
  int rc = fa_forward_ex(&p, nullptr);
  if (rc) { printf("tcgen05 launch failed: %s %s\n", fa_strerror(rc), fa_last_cuda_error()); return 3; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    uint32_t wd[4] = {0, 0, 0, 0};
    fa_watchdog_info(wd);
    printf("tcgen05 kernel failed: %s (watchdog tag %u block %u thread %u parity %u)\n", cudaGetErrorString(e), wd[0], wd[1], wd[2], wd[3]);
    return 3;
  }

  auto download = [&](void* dptr, std::vector<float>& h) -> int {
    h.resize(n_el);
    if (!bf16) { CK(cudaMemcpy(h.data(), dptr, n_el * 4, cudaMemcpyDeviceToHost)); }
    else {
      std::vector<uint16_t> t(n_el);
      CK(cudaMemcpy(t.data(), dptr, n_el * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n_el; ++i) {
        if (f16) { __half x; memcpy(&x, &t[i], 2); h[i] = __half2float(x); }
        else { __nv_bfloat16 x; memcpy(&x, &t[i], 2); h[i] = __bfloat162float(x); }
      }
    }
    return 0;
  };
  std::vector<float> o_tc, o_simt, lse_tc(BH * N), lse_simt(BH * N);
  if (download(do_tc, o_tc)) return 2;
  CK(cudaMemcpy(lse_tc.data(), dlse_tc, BH * N * 4, cudaMemcpyDeviceToHost));

  double tc_vs_simt = -1, lse_tc_vs_simt = -1;
  if (check_simt) {
    p.o = do_simt; p.lse = dlse_simt; p.impl = FA_IMPL_SIMT;
    rc = fa_forward_ex(&p, nullptr);
    if (rc) { printf("simt launch failed: %s %s\n", fa_strerror(rc), fa_last_cuda_error()); return 3; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("simt kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
    if (download(do_simt, o_simt)) return 2;
    CK(cudaMemcpy(lse_simt.data(), dlse_simt, BH * N * 4, cudaMemcpyDeviceToHost));
    tc_vs_simt = 0; lse_tc_vs_simt = 0;
    for (size_t i = 0; i < n_el; ++i) {
      const double df = fabs((double)o_tc[i] - o_simt[i]);
      if (df > tc_vs_simt || df != df) tc_vs_simt = df;
    }
    for (int64_t i = 0; i < BH * N; ++i) {
      const double df = fabs((double)lse_tc[i] - lse_simt[i]);
      if (df > lse_tc_vs_simt || df != df) lse_tc_vs_simt = df;
    }
  }

  // fp64 host evaluation on sampled rows
  double err_tc = 0, err_simt = 0, err_lse = 0;
  const int n_samples = 48;
  std::vector<double> sc(N);
  for (int sidx = 0; sidx < n_samples; ++sidx) {
    const int64_t bh = (int64_t)(urand() * BH) % BH;
    int64_t row = (int64_t)(urand() * N) % N;
    if (sidx == 0) row = 0;
    if (sidx == 1) row = N - 1;
    if (sidx == 2) row = std::min<int64_t>(N - 1, 127);
    if (sidx == 3) row = std::min<int64_t>(N - 1, 128);
    const float* q = &hq[(bh * N + row) * d];
    const int64_t last = causal ? row : N - 1;
    double mx = -1e300;
    for (int64_t j = 0; j <= last; ++j) {
      const float* kk = &hk[(bh * N + j) * d];
      double s = 0;
      for (int c = 0; c < d; ++c) s += (double)q[c] * kk[c];
      sc[j] = s * scale;
      mx = std::max(mx, sc[j]);
    }
    double l = 0;
    for (int64_t j = 0; j <= last; ++j) { sc[j] = exp(sc[j] - mx); l += sc[j]; }
    for (int c = 0; c < d; ++c) {
      double o = 0;
      for (int64_t j = 0; j <= last; ++j) o += sc[j] * hv[(bh * N + j) * d + c];
      o /= l;
      err_tc = std::max(err_tc, fabs(o - o_tc[(bh * N + row) * d + c]));
      if (o_tc[(bh * N + row) * d + c] != o_tc[(bh * N + row) * d + c]) err_tc = 1e30;
      if (check_simt) err_simt = std::max(err_simt, fabs(o - o_simt[(bh * N + row) * d + c]));
    }
    err_lse = std::max(err_lse, fabs(mx + log(l) - lse_tc[bh * N + row]));
  }

  // timing (tcgen05 path), L2 flushed between repetitions
  p.o = do_tc; p.lse = nullptr; p.impl = FA_IMPL_TCGEN05;
  if (const char* ti = getenv("FA_CHECK_TIME_IMPL")) p.impl = atoi(ti);   // 2 = time the CUDA-core kernel instead
  void* flush;
  const size_t flush_bytes = 256u << 20;
  CK(cudaMalloc(&flush, flush_bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) fa_forward_ex(&p, nullptr);
  CK(cudaDeviceSynchronize());
  std::vector<float> ms(reps);
  for (int i = 0; i < reps; ++i) {
    CK(cudaMemsetAsync(flush, i, flush_bytes, nullptr));
    CK(cudaEventRecord(e0, nullptr));
    fa_forward_ex(&p, nullptr);
    CK(cudaEventRecord(e1, nullptr));
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      uint32_t wd[4] = {0, 0, 0, 0};
      fa_watchdog_info(wd);
      printf("timed launch %d failed (watchdog tag %u block %u thread %u parity %u)\n", i, wd[0], wd[1], wd[2], wd[3]);
      return 3;
    }
    CK(cudaEventElapsedTime(&ms[i], e0, e1));
  }
  std::sort(ms.begin(), ms.end());
  const double med = reps ? ms[reps / 2] : 0, mn = reps ? ms[0] : 0;
  const double flops = 4.0 * BH * (double)N * N * d * (causal ? 0.5 : 1.0);
  const double bytes = 4.0 * BH * (double)N * d * es;
  printf("{\"check\": \"%s d=%d BH=%lld N=%lld causal=%d scale=%g\", \"err_tc_vs_fp64\": %.3e, \"err_simt_vs_fp64\": %.3e, "
         "\"tc_vs_simt\": %.3e, \"lse_err\": %.3e, \"lse_tc_vs_simt\": %.3e, \"ms_median\": %.4f, \"ms_min\": %.4f, "
         "\"tflops_median\": %.1f, \"gbs_median\": %.1f}\n",
         argv[1], d, (long long)BH, (long long)N, causal, scale, err_tc, err_simt, tc_vs_simt, err_lse, lse_tc_vs_simt, med, mn,
         med > 0 ? flops / med * 1e-9 : 0.0, med > 0 ? bytes / med * 1e-6 : 0.0);
  return 0;
}
