// softmax_probe.cu — microbenchmark of the softmax inner step of the forward kernel (tcgen05.ld of an S row ->
// row max -> exp2 -> pack -> tcgen05.st of P), isolated from the MMA pipeline.  Question it answers: how many
// cycles does ONE 128x128 S tile cost the softmax role when
//   W=1: one thread owns a whole 128-key row (one warp per SM sub-partition, as shipped up to v6), or
//   W=2: two threads share a row, 64 keys each (warps w and w+4 sit on the same sub-partition and the same TMEM
//        lanes; the row max is exchanged through SMEM),
// and what the 3-input max and the FMA-pipe polynomial exp2 are worth in either layout.
//   ./softmax_probe [iters]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../csrc/ptx.cuh"

using namespace fa;

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// kCols keys of one row, in registers: exp2(s*c - mc) in place, partial sums into l0..l3.  kPoly = how many of
// every 8 elements use the polynomial (0, 2, 4).
template <bool kTF32, int kPoly>
__device__ __forceinline__ void exp_block(float* s, int i0, int n, float c, float neg_mc, float& l0, float& l1, float& l2, float& l3) {
#pragma unroll
  for (int i = i0; i < i0 + n; i += 4) {
    const bool poly01 = (kPoly >= 2) && ((i & 4) == 0);
    const bool poly23 = (kPoly >= 4) && ((i & 4) == 0);
    if constexpr (!kTF32) {
      float2 a01 = ffma2(make_float2(s[i], s[i + 1]), make_float2(c, c), make_float2(neg_mc, neg_mc));
      float2 a23 = ffma2(make_float2(s[i + 2], s[i + 3]), make_float2(c, c), make_float2(neg_mc, neg_mc));
      if (poly01) { a01 = exp2_poly2(a01); s[i] = a01.x; s[i + 1] = a01.y; } else { s[i] = ex2(a01.x); s[i + 1] = ex2(a01.y); }
      if (poly23) { a23 = exp2_poly2(a23); s[i + 2] = a23.x; s[i + 3] = a23.y; } else { s[i + 2] = ex2(a23.x); s[i + 3] = ex2(a23.y); }
      const float2 s01 = fadd2(make_float2(l0, l1), make_float2(s[i], s[i + 1]));
      const float2 s23 = fadd2(make_float2(l2, l3), make_float2(s[i + 2], s[i + 3]));
      l0 = s01.x; l1 = s01.y; l2 = s23.x; l3 = s23.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool poly = k < 2 ? poly01 : poly23;
        float v = fmaf(s[i + k], c, neg_mc);
        v = poly ? exp2_poly(v) : ex2(v);
        s[i + k] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      }
      l0 += s[i]; l1 += s[i + 1]; l2 += s[i + 2]; l3 += s[i + 3];
    }
  }
}

// kW = threads per row (1 or 2); kMax3: 3-input max; kQuarters: deliver P in 32-key pieces (tcgen05.st + wait per piece)
template <bool kTF32, int kW, bool kMax3, int kPoly>
__global__ void __launch_bounds__(128 * kW, 1) probe(float* out, long long* cycles, int iters, float c) {
  __shared__ uint32_t s_tptr[4];
  constexpr int kCols = 128 / kW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2;   // which key half of the row (kW == 2)
  if (warp == 0) {
    tmem_alloc(smem_u32(s_tptr), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tptr[0];
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_base + 128 + half * kCols;       // source S (never overwritten)
  const uint32_t tP = tmem_base + lane_base + (kTF32 ? half * kCols : half * (kCols / 2));
  // fill S with N(0,1)-ish values
  {
    uint32_t v[32];
    uint32_t x = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
    for (int q = 0; q < kCols / 32; ++q) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        x = x * 1664525u + 1013904223u;
        v[i] = __float_as_uint(((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 6.0f);
      }
      tmem_st32(tS + q * 32, v);
    }
    tc_wait_st();
  }
  __syncthreads();
  float m = -INFINITY, l = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float s[kCols];
#pragma unroll
    for (int q = 0; q < kCols / 32; ++q) tmem_ld32(tS + q * 32, reinterpret_cast<uint32_t*>(&s[q * 32]));
    tc_wait_ld();
    float mx;
    if constexpr (kMax3) {
      float a = -INFINITY, b = -INFINITY;
#pragma unroll
      for (int i = 0; i < kCols; i += 4) {
        a = max3(a, s[i], s[i + 1]);
        b = max3(b, s[i + 2], s[i + 3]);
      }
      mx = fmaxf(a, b);
    } else {
      float a = -INFINITY, b = -INFINITY, d = -INFINITY, e = -INFINITY;
#pragma unroll
      for (int i = 0; i < kCols; i += 4) {
        a = fmaxf(a, s[i]); b = fmaxf(b, s[i + 1]); d = fmaxf(d, s[i + 2]); e = fmaxf(e, s[i + 3]);
      }
      mx = fmaxf(fmaxf(a, b), fmaxf(d, e));
    }
    float m_new = fmaxf(m, mx);
    if constexpr (kW == 2) {
      // exchange the partial row max with the thread that owns the other 64 keys of this row: each half writes its
      // own slot (double-buffered over iterations), pair barrier (warps w and w+4), read the partner's
      __shared__ float s_x[2][2][128];
      const int r = (warp & 3) * 32 + lane;
      s_x[it & 1][half][r] = mx;
      named_bar_sync(1 + (warp & 3), 64);
      m_new = fmaxf(m_new, s_x[it & 1][half ^ 1][r]);
    }
    const bool need = (m_new - m) * c > 8.0f;
    if (__any_sync(0xffffffffu, need)) {
      const float m_use = need ? m_new : m;
      l *= need ? ex2((m - m_use) * c) : 1.0f;
      m = m_use;
    }
    const float neg_mc = -((m == -INFINITY) ? 0.f : m) * c;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    constexpr int kPiece = kW == 1 ? 64 : 32;   // P is handed over in two pieces per thread
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      exp_block<kTF32, kPoly>(s, h * kPiece, kPiece, c, neg_mc, l0, l1, l2, l3);
      if constexpr (kTF32) {
#pragma unroll
        for (int q = 0; q < kPiece / 32; ++q) tmem_st32(tP + h * kPiece + q * 32, reinterpret_cast<uint32_t*>(&s[h * kPiece + q * 32]));
      } else {
        uint32_t pk[kPiece / 2];
#pragma unroll
        for (int i = 0; i < kPiece / 2; ++i) pk[i] = pack_bf16x2(s[h * kPiece + 2 * i], s[h * kPiece + 2 * i + 1]);
        if constexpr (kPiece == 64) tmem_st32(tP + h * 32, pk);
        else tmem_st16(tP + h * 16, pk);
      }
      tc_wait_st();
      tc_fence_before();
    }
    l += (l0 + l1) + (l2 + l3);
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * (4 * kW) + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// G independent one-thread-per-row groups (the two tile slots of the forward kernel, free-running) plus kSpin warps that
// wait on an mbarrier which only completes when the workers are done (what the idle slot / producer / MMA warps do).
template <int kG, int kSpin, int kPoly>
__global__ void __launch_bounds__(128 * kG + 32 * kSpin, 1) probe_groups(float* out, long long* cycles, int iters, float c) {
  __shared__ uint32_t s_tptr[4];
  __shared__ __align__(8) unsigned long long s_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar), 4 * kG);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(s_tptr), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tptr[0];
  if (warp >= 4 * kG) {
    mbar_wait(smem_u32(&s_bar), 0, 99);
  } else {
    const int g = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + 256 + g * 128;
    const uint32_t tP = tmem_base + lane_base + g * 128;
    {
      uint32_t v[32];
      uint32_t x = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x = x * 1664525u + 1013904223u;
          v[i] = __float_as_uint(((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 6.0f);
        }
        tmem_st32(tS + q * 32, v);
      }
      tc_wait_st();
    }
    named_bar_sync(1, 128 * kG);
    if (g == 1) __nanosleep(400);   // start the second slot out of phase
    float m = -INFINITY, l = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      float s[128];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld32(tS + q * 32, reinterpret_cast<uint32_t*>(&s[q * 32]));
      tc_wait_ld();
      float a = -INFINITY, b = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        a = max3(a, s[i], s[i + 1]);
        b = max3(b, s[i + 2], s[i + 3]);
      }
      const float m_new = fmaxf(m, fmaxf(a, b));
      const bool need = (m_new - m) * c > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float m_use = need ? m_new : m;
        l *= need ? ex2((m - m_use) * c) : 1.0f;
        m = m_use;
      }
      const float neg_mc = -((m == -INFINITY) ? 0.f : m) * c;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        exp_block<false, kPoly>(s, h * 64, 64, c, neg_mc, l0, l1, l2, l3);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pk[i] = pack_bf16x2(s[h * 64 + 2 * i], s[h * 64 + 2 * i + 1]);
        tmem_st32(tP + h * 32, pk);
        tc_wait_st();
        tc_fence_before();
      }
      l += (l0 + l1) + (l2 + l3);
    }
    const long long t1 = clock64();
    if (lane == 0) {
      cycles[blockIdx.x * 8 + warp] = t1 - t0;
      mbar_arrive(smem_u32(&s_bar));
    }
    out[blockIdx.x * 256 + threadIdx.x] = l + m;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int kG, int kSpin, int kPoly>
void run_groups(const char* name, int iters) {
  const int grid = 148;
  float* out;
  long long* cyc;
  cudaMalloc(&out, grid * 256 * sizeof(float));
  cudaMalloc(&cyc, grid * 8 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) probe_groups<kG, kSpin, kPoly><<<grid, 128 * kG + 32 * kSpin>>>(out, cyc, iters, 0.1275f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<long long> h(grid * 8);
  cudaMemcpy(h.data(), cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  int n = 0;
  for (int b = 0; b < grid; ++b)
    for (int w = 0; w < 4 * kG; ++w) { sum += h[b * 8 + w]; ++n; }
  printf("{\"probe\": \"%s\", \"groups\": %d, \"spinner_warps\": %d, \"cycles_per_tile_per_group\": %.1f, \"cycles_per_tile_aggregate\": %.1f}\n",
         name, kG, kSpin, sum / n / iters, sum / n / iters / kG);
  cudaFree(out);
  cudaFree(cyc);
}

template <bool kTF32, int kW, bool kMax3, int kPoly>
void run(const char* name, int iters) {
  const int grid = 148;
  float* out;
  long long* cyc;
  cudaMalloc(&out, grid * 256 * sizeof(float));
  cudaMalloc(&cyc, grid * 8 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) probe<kTF32, kW, kMax3, kPoly><<<grid, 128 * kW>>>(out, cyc, iters, 0.1275f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<long long> h(grid * 4 * kW);
  cudaMemcpy(h.data(), cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  long long mn = 1ll << 60, mxv = 0;
  for (long long v : h) { sum += v; mn = v < mn ? v : mn; mxv = v > mxv ? v : mxv; }
  printf("{\"probe\": \"%s\", \"threads_per_row\": %d, \"cycles_per_tile_avg\": %.1f, \"min\": %.1f, \"max\": %.1f, \"cycles_per_elem_row\": %.2f}\n",
         name, kW, sum / h.size() / iters, (double)mn / iters, (double)mxv / iters, sum / h.size() / iters / 128.0);
  cudaFree(out);
  cudaFree(cyc);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  run<false, 1, false, 0>("bf16 W1 max2 poly0 (shipped)", iters);
  run<false, 1, true, 0>("bf16 W1 max3 poly0", iters);
  run<false, 1, true, 2>("bf16 W1 max3 poly2", iters);
  run<false, 2, false, 0>("bf16 W2 max2 poly0", iters);
  run<false, 2, true, 0>("bf16 W2 max3 poly0", iters);
  run<false, 2, true, 2>("bf16 W2 max3 poly2", iters);
  run<false, 2, true, 4>("bf16 W2 max3 poly4", iters);
  run<true, 1, false, 0>("tf32 W1 max2 poly0 (shipped)", iters);
  run<true, 1, true, 0>("tf32 W1 max3 poly0", iters);
  run<true, 2, true, 0>("tf32 W2 max3 poly0", iters);
  run<true, 2, true, 2>("tf32 W2 max3 poly2", iters);
  run_groups<1, 0, 0>("bf16 G1 spin0 poly0", iters);
  run_groups<1, 4, 0>("bf16 G1 spin4 poly0", iters);
  run_groups<1, 6, 0>("bf16 G1 spin6 poly0", iters);
  run_groups<2, 0, 0>("bf16 G2 spin0 poly0", iters);
  run_groups<2, 2, 0>("bf16 G2 spin2 poly0", iters);
  run_groups<2, 0, 2>("bf16 G2 spin0 poly2", iters);
  run_groups<2, 0, 4>("bf16 G2 spin0 poly4", iters);
  return 0;
}
