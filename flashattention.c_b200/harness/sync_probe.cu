// sync_probe.cu — cost (SM cycles) of the synchronisation primitives on the MMA warp's path, measured on completed
// barriers (i.e. pure overhead): mbarrier try_wait / test_wait, tcgen05.fence::after_thread_sync, elect.sync,
// tcgen05.commit with an idle tensor pipe.   ./sync_probe
#include <cuda_runtime.h>

#include <cstdio>

#include "../csrc/ptx.cuh"
using namespace fa;

__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(32, 1) probe(long long* out) {
  __shared__ __align__(8) unsigned long long bars[8];
  __shared__ uint32_t s_tptr[4];
  const uint32_t b0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(b0 + 8 * i, 1);
    fence_mbar_init();
  }
  tmem_alloc(smem_u32(s_tptr), 32);
  tmem_relinquish();
  __syncwarp();
  if (threadIdx.x == 0)
    for (int i = 0; i < 4; ++i) mbar_arrive(b0 + 8 * i);   // phases 0 of bars 0..3 complete
  __syncwarp();
  constexpr int N = 64;
  long long t[8];
  int acc = 0;
  t[0] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) acc += mbar_try_wait(b0 + 8 * (i & 3), 0);
  t[1] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) acc += mbar_test_wait(b0 + 8 * (i & 3), 0);
  t[2] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { mbar_wait(b0 + 8 * (i & 3), 0, 1); tc_fence_after(); }
  t[3] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { acc += elect_one_sync(); __syncwarp(); }
  t[4] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { tc_fence_after(); }
  t[5] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { if (elect_one_sync()) tc_commit(b0 + 8 * 4); __syncwarp(); }
  t[6] = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { if (threadIdx.x == 0) mbar_arrive(b0 + 8 * 5); __syncwarp(); }
  t[7] = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) out[i] = (t[i + 1] - t[i]) / N;
    out[7] = acc;
  }
  __syncwarp();
  tmem_dealloc(s_tptr[0], 32);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * sizeof(long long));
  probe<<<1, 32>>>(d);
  probe<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[7] = {"try_wait(completed)", "test_wait(completed)", "mbar_wait+tcgen05.fence::after", "elect.sync+syncwarp",
                          "tcgen05.fence::after", "elect+tcgen05.commit (idle pipe)", "mbarrier.arrive (lane 0)+syncwarp"};
  for (int i = 0; i < 7; ++i) printf("{\"primitive\": \"%s\", \"cycles_per_op\": %lld}\n", names[i], h[i]);
  return 0;
}
