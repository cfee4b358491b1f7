// llmc_main.cu — the kernel-6 branch of the reference's llm.c harness (src/llm.c/attention_forward.cu:1214-1287), restated
// over the exported C symbol `attention_forward(kernel_num, out, vaccum, qkvr, preatt, att, inp, B, T, C, NH, block_size)` of
// libfa_b200.so.  TEST INFRASTRUCTURE: the CPU side of the comparison is the oracle's restatement of attention_forward_cpu
// (oracle/fa_oracle.c, pinned bit-for-bit to the reference's own function by tests/test_oracle.py), linked from
// oracle/_build/libfa_oracle.so.
//
// Same as the reference harness: srand(0) + rand() inputs in [-1, 1) (common.h:46-52), B=6 T=4096 C=768 NH=12, the five block
// sizes, `out` compared element by element at an absolute 1e-4 (line 1262; common.h:79-106: the first five pairs are printed,
// the process exits once ten elements are off), then the averaged cudaEvent timing loop (common.h:108-124).  Not reproduced:
// the 2 x 4.8 GB preatt/att buffers, which kernel 6 never touches (lines 1264-1274), and kernels 1-5.
//
//   llmc_main [kernel_num=6] [B=6] [T=4096] [repeats=100]      exit 0 only if EVERY element of every run is within 1e-4
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/fa_b200.h"

extern "C" void fa_oracle_llmc_cpu(float* out, const float* inp, int B, int T, int C, int NH);

#define CUDA_OK(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      printf("[CUDA ERROR] at file %s:%d:\n%s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
      exit(EXIT_FAILURE);                                                                          \
    }                                                                                              \
  } while (0)

// returns the number of elements beyond the tolerance; exits at the tenth, like the reference's validate_result
static size_t validate(const float* device_result, const float* cpu_reference, const char* name, size_t n, float tolerance,
                       double* max_err) {
  float* got = (float*)malloc(n * sizeof(float));
  CUDA_OK(cudaMemcpy(got, device_result, n * sizeof(float), cudaMemcpyDeviceToHost));
  size_t faults = 0;
  for (size_t i = 0; i < n; i++) {
    if (i < 5) printf("%f %f\n", cpu_reference[i], got[i]);
    const double err = fabs((double)cpu_reference[i] - (double)got[i]);
    if (!std::isnan(cpu_reference[i]) && err > *max_err) *max_err = err;
    if (!(err <= tolerance) && !std::isnan(cpu_reference[i])) {
      printf("Mismatch of %s at %zu: CPU_ref: %f vs GPU: %f\n", name, i, cpu_reference[i], got[i]);
      if (++faults >= 10) {
        free(got);
        exit(EXIT_FAILURE);
      }
    }
  }
  free(got);
  return faults;
}

int main(int argc, char** argv) {
  srand(0);
  const int kernel_num = argc > 1 ? atoi(argv[1]) : 6;
  const int B = argc > 2 ? atoi(argv[2]) : 6;
  const int T = argc > 3 ? atoi(argv[3]) : 4096;
  const int repeats = argc > 4 ? atoi(argv[4]) : 100;
  const int C = 768, NH = 12;
  CUDA_OK(cudaSetDevice(0));

  const size_t n_out = (size_t)B * T * C, n_inp = 3 * n_out;
  float* out = (float*)malloc(n_out * sizeof(float));
  float* inp = (float*)malloc(n_inp * sizeof(float));
  for (size_t i = 0; i < n_inp; i++) inp[i] = ((float)rand() / RAND_MAX) * 2.0 - 1.0;
  float *d_out, *d_vaccum, *d_qkvr, *d_inp;
  CUDA_OK(cudaMalloc(&d_out, n_out * sizeof(float)));
  CUDA_OK(cudaMalloc(&d_vaccum, n_out * sizeof(float)));
  CUDA_OK(cudaMalloc(&d_qkvr, n_inp * sizeof(float)));
  CUDA_OK(cudaMalloc(&d_inp, n_inp * sizeof(float)));
  CUDA_OK(cudaMemcpy(d_inp, inp, n_inp * sizeof(float), cudaMemcpyHostToDevice));
  float* d_preatt = nullptr;   // (B, NH, T, T) in the reference; kernel 6 neither reads nor writes them
  float* d_att = nullptr;

  printf("Using kernel %d\n", kernel_num);
  const int block_sizes[] = {32, 64, 128, 256, 512};
  fa_oracle_llmc_cpu(out, inp, B, T, C, NH);
  size_t faults = 0;
  double max_err = 0.0;
  for (int block_size : block_sizes) {
    printf("Checking block size %d.\n", block_size);
    CUDA_OK(cudaMemset(d_out, 0xff, n_out * sizeof(float)));   // NaN pattern: a row the call does not write cannot pass
    attention_forward(kernel_num, d_out, d_vaccum, d_qkvr, d_preatt, d_att, d_inp, B, T, C, NH, block_size);
    faults += validate(d_out, out, "out", n_out, 1e-4f, &max_err);
  }
  printf("max |out - cpu| = %.3g over %zu elements x %zu runs, %zu beyond 1e-4\n", max_err, n_out, sizeof(block_sizes) / sizeof(int), faults);
  if (faults) return 2;
  printf("All results match. Starting benchmarks.\n\n");

  for (int block_size : block_sizes) {
    cudaEvent_t start, stop;
    CUDA_OK(cudaEventCreate(&start));
    CUDA_OK(cudaEventCreate(&stop));
    CUDA_OK(cudaEventRecord(start, nullptr));
    for (int i = 0; i < repeats; i++)
      attention_forward(kernel_num, d_out, d_vaccum, d_qkvr, d_preatt, d_att, d_inp, B, T, C, NH, block_size);
    CUDA_OK(cudaEventRecord(stop, nullptr));
    CUDA_OK(cudaEventSynchronize(stop));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, start, stop));
    printf("block_size %4d | time %f ms\n", block_size, ms / repeats);
    CUDA_OK(cudaEventDestroy(start));
    CUDA_OK(cudaEventDestroy(stop));
  }
  free(out);
  free(inp);
  CUDA_OK(cudaFree(d_out));
  CUDA_OK(cudaFree(d_vaccum));
  CUDA_OK(cudaFree(d_qkvr));
  CUDA_OK(cudaFree(d_inp));
  return 0;
}
