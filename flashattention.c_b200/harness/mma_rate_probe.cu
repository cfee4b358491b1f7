// mma_rate_probe.cu — how fast does one thread get tcgen05.mma through the tensor pipe, by operand source and shape?
// One CTA, one issuing thread, R MMAs back to back into one accumulator, one commit, clock64 around it.
//   SS: A and B from shared memory (K-major, SWIZZLE_128B: what S = Q K^T uses)      TS: A from tensor memory (what P V uses)
// The question it answers: is kind::tf32 S = Q K^T (32-byte K-slices of 128-byte swizzled rows, both operands from SMEM)
// bound by the operand fetch rather than by the pipe — and would Q as a TMEM operand lift that bound?
#include <cstdio>
#include <cstdlib>

#include "../csrc/ptx.cuh"

using namespace fa;

template <bool kTF32, bool kFromTmem, int kN>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps, int inner) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 4 * 16384, bar = base + 8 * 16384, tptr = bar + 16;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  // operands: small finite numbers
  for (uint32_t i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) st_shared_b32(base + 4 * i, 0x3c003c00u);
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ld_shared_b32(tptr);
  if (threadIdx.x == 0) {
    constexpr uint32_t fmt = kTF32 ? 2u : 1u;
    constexpr uint32_t idesc = make_idesc(fmt, 0, 128, kN);
    constexpr uint64_t hi = make_sdesc_hi_sw128(16, 1024);
    constexpr int kSteps = kTF32 ? 16 : 8;   // k-steps in a 128-row x (2 chunks of 128 B) operand tile, as in the kernel
    long long best = 1ll << 60;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int it = 0; it < inner; ++it) {
#pragma unroll
        for (int kk = 0; kk < kSteps; ++kk) {
          const uint32_t off16 = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if constexpr (kFromTmem) mma_ts<kTF32>(tmem, tmem + 256 + kk * 8, sdesc_at(hi, sB) + off16, idesc, 1u);
          else mma_ss<kTF32>(tmem, sdesc_at(hi, sA) + off16, sdesc_at(hi, sB) + off16, idesc, 1u);
        }
      }
      tc_commit(bar);
      mbar_wait(bar, parity, 99);
      parity ^= 1;
      const long long dt = clock64() - t0;
      best = dt < best ? dt : best;
    }
    out[0] = best;
    out[1] = (long long)inner * kSteps;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// ---- tensor-memory contention: the same MMA stream while `busy_warps` other warps of the CTA stream tcgen05.ld / tcgen05.st
// over other TMEM columns (what the softmax warps do).  Does concurrent softmax traffic slow the MMAs down?
template <bool kTF32, bool kFromTmem, int kN>
__global__ void __launch_bounds__(288, 1) probe_contended(long long* out, int reps, int inner, int busy_warps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 4 * 16384, bar = base + 8 * 16384, tptr = bar + 16, flag = bar + 32;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    st_shared_b32(flag, 1u);
    fence_mbar_init();
  }
  for (uint32_t i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) st_shared_b32(base + 4 * i, 0x3c003c00u);
  fence_proxy_async_smem();
  if (warp == 8) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ld_shared_b32(tptr);
  if (warp < 8) {
    if (warp < busy_warps) {
      // lanes 32 * (warp % 4) .. + 31, columns [256, 384) (warps 0-3) or [384, 512) (warps 4-7): away from the MMA's D and A columns
      const uint32_t t = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 128;
      uint32_t r[32];
      for (int spin = 0; spin < 200000 && ld_shared_b32(flag) != 0u; ++spin) {   // bounded: a probe must never hang the GPU
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tmem_ld32(t + q * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] += 1u;
          tmem_st32(t + q * 32, r);
        }
        tc_wait_st();
      }
    }
  } else if (threadIdx.x == 256) {
    constexpr uint32_t fmt = kTF32 ? 2u : 1u;
    constexpr uint32_t idesc = make_idesc(fmt, 0, 128, kN);
    constexpr uint64_t hi = make_sdesc_hi_sw128(16, 1024);
    constexpr int kSteps = kTF32 ? 16 : 8;
    long long best = 1ll << 60;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int it = 0; it < inner; ++it) {
#pragma unroll
        for (int kk = 0; kk < kSteps; ++kk) {
          const uint32_t off16 = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if constexpr (kFromTmem) mma_ts<kTF32>(tmem, tmem + 128 + kk * 8, sdesc_at(hi, sB) + off16, idesc, 1u);
          else mma_ss<kTF32>(tmem, sdesc_at(hi, sA) + off16, sdesc_at(hi, sB) + off16, idesc, 1u);
        }
      }
      tc_commit(bar);
      mbar_wait(bar, parity, 97);
      parity ^= 1;
      const long long dt = clock64() - t0;
      best = dt < best ? dt : best;
    }
    out[0] = best;
    out[1] = (long long)inner * kSteps;
    st_shared_b32(flag, 0u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

template <bool kTF32, bool kFromTmem, int kN>
void run_contended(const char* name, int busy_warps) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  auto k = probe_contended<kTF32, kFromTmem, kN>;
  const int smem = 8 * 16384 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 288, smem>>>(d, 5, 8, busy_warps);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  printf("{\"probe\": \"%s, %d warps streaming tcgen05.ld/st\", \"mmas\": %lld, \"cycles\": %lld, \"cycles_per_mma\": %.1f}\n", name, busy_warps, h[1],
         h[0], (double)h[0] / h[1]);
  fflush(stdout);
  cudaFree(d);
}

// ---- the same with a CTA pair: one leader thread issues tcgen05.mma.cta_group::2 (M = 256: 128 rows in each CTA's tensor memory,
// each CTA holds half of B's N extent in its shared memory), the commit is multicast to a barrier in both CTAs.
FA_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
FA_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool kTF32, bool kFromTmem, int kN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair(long long* out, int reps, int inner) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 4 * 16384, bar = base + 8 * 16384, tptr = bar + 16;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (uint32_t i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) st_shared_b32(base + 4 * i, 0x3c003c00u);
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = ld_shared_b32(tptr);
  if (threadIdx.x == 0) {
    constexpr uint32_t fmt = kTF32 ? 2u : 1u;
    constexpr uint32_t idesc = make_idesc(fmt, 0, 256, kN);
    constexpr uint64_t hi = make_sdesc_hi_sw128(16, 1024);
    constexpr int kSteps = kTF32 ? 16 : 8;
    long long best = 1ll << 60;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (rank == 0) {
        for (int it = 0; it < inner; ++it) {
#pragma unroll
          for (int kk = 0; kk < kSteps; ++kk) {
            const uint32_t off16 = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
            const uint64_t bd = sdesc_at(hi, sB) + off16, ad = sdesc_at(hi, sA) + off16;
            const uint32_t at = tmem + 256 + kk * 8;
            if constexpr (kFromTmem && kTF32)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(at), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else if constexpr (kFromTmem)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(at), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else if constexpr (kTF32)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"((uint16_t)3) : "memory");
      }
      mbar_wait(bar, parity, 98);
      parity ^= 1;
      const long long dt = clock64() - t0;
      best = dt < best ? dt : best;
    }
    if (rank == 0) {
      out[0] = best;
      out[1] = (long long)inner * kSteps;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <bool kTF32, bool kFromTmem, int kN>
void run_pair(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  auto k = probe_pair<kTF32, kFromTmem, kN>;
  const int smem = 8 * 16384 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<2, 128, smem>>>(d, 5, 8);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  printf("{\"probe\": \"%s\", \"mmas\": %lld, \"cycles\": %lld, \"cycles_per_mma\": %.1f}\n", name, h[1], h[0], (double)h[0] / h[1]);
  fflush(stdout);
  cudaFree(d);
}

template <bool kTF32, bool kFromTmem, int kN>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  auto k = probe<kTF32, kFromTmem, kN>;
  const int smem = 8 * 16384 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(d, 5, 8);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  printf("{\"probe\": \"%s\", \"mmas\": %lld, \"cycles\": %lld, \"cycles_per_mma\": %.1f}\n", name, h[1], h[0], (double)h[0] / h[1]);
  fflush(stdout);
  cudaFree(d);
}

// ---- the backward kernel's MMA stream, alone on the SM: per step  TS (acc0 += P^T dO)  SS (S)  TS (acc1 += dS^T Q)  SS (dP)
// (fa_bwd_sm100.cuh), one issuing thread, nothing else running.  order 0: as the kernel issues them (four switches between
// SMEM- and TMEM-sourced A operands per step); 1: TS TS SS SS (two switches); 2: the SS groups only; 3: the TS groups only.
// What does a step cost the tensor pipe by itself, and what does a switch cost?
template <int kD>
__global__ void __launch_bounds__(128, 1) probe_bwd_pattern(long long* out, int reps, int inner, int order) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sR = base, sT = base + 4 * 16384, bar = base + 8 * 16384, tptr = bar + 16;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (uint32_t i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) st_shared_b32(base + 4 * i, 0x3c003c00u);
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ld_shared_b32(tptr);
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc_sd = make_idesc(1u, 0, 128, 128);
    constexpr uint32_t idesc_acc = make_idesc(1u, 1, 128, kD);
    constexpr uint64_t hi_k = make_sdesc_hi_sw128(16, 1024);
    constexpr uint64_t hi_mn = make_sdesc_hi_sw128(16384, 1024);
    constexpr int kStepsD = kD / 16;
    auto ss = [&](uint32_t d) {
#pragma unroll
      for (int kk = 0; kk < kStepsD; ++kk) {
        const uint32_t off16 = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
        mma_ss<false>(d, sdesc_at(hi_k, sR) + off16, sdesc_at(hi_k, sT) + off16, idesc_sd, kk > 0 ? 1u : 0u);
      }
    };
    auto ts = [&](uint32_t acc, uint32_t a) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        mma_ts<false>(acc, a + static_cast<uint32_t>((ks >> 2) * 64 + (ks & 3) * 8), sdesc_at(hi_mn, sT) + static_cast<uint32_t>(ks * 128), idesc_acc, 1u);
    };
    long long best = 1ll << 60;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int it = 0; it < inner; ++it) {
        if (order == 0) { ts(tmem + 256, tmem); ss(tmem); ts(tmem + 256 + kD, tmem + 128); ss(tmem + 128); }
        else if (order == 1) { ts(tmem + 256, tmem); ts(tmem + 256 + kD, tmem + 128); ss(tmem); ss(tmem + 128); }
        else if (order == 2) { ss(tmem); ss(tmem + 128); }
        else { ts(tmem + 256, tmem); ts(tmem + 256 + kD, tmem + 128); }
      }
      tc_commit(bar);
      mbar_wait(bar, parity, 99);
      parity ^= 1;
      const long long dt = clock64() - t0;
      best = dt < best ? dt : best;
    }
    out[0] = best;
    out[1] = inner;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
template <int kD>
void run_bwd_pattern(int order, const char* what) {
  long long* d;
  cudaMalloc(&d, 16);
  auto k = probe_bwd_pattern<kD>;
  const int smem = 8 * 16384 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(d, 5, 8, order);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("bwd pattern: %s\n", cudaGetErrorString(e)); exit(1); }
  printf("{\"probe\": \"backward MMA stream d=%d, %s\", \"steps\": %lld, \"cycles\": %lld, \"cycles_per_step\": %.1f}\n", kD, what, h[1], h[0],
         (double)h[0] / h[1]);
  fflush(stdout);
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'b') {   // only the backward-pattern probes
    run_bwd_pattern<128>(0, "TS SS TS SS (as issued: 16 TS + 16 SS MMAs per step)");
    run_bwd_pattern<128>(1, "TS TS SS SS");
    run_bwd_pattern<128>(2, "SS groups only (16 MMAs)");
    run_bwd_pattern<128>(3, "TS groups only (16 MMAs)");
    run_bwd_pattern<64>(0, "TS SS TS SS (as issued: 16 TS N=64 + 8 SS MMAs per step)");
    run_bwd_pattern<64>(1, "TS TS SS SS");
    run_bwd_pattern<64>(2, "SS groups only (8 MMAs)");
    run_bwd_pattern<64>(3, "TS groups only (16 MMAs, N = 64)");
    return 0;
  }
  run<true, false, 128>("tf32 SS 128x128x8  (S = Q K^T as shipped)");
  run<true, true, 128>("tf32 TS 128x128x8  (Q from TMEM)");
  run<true, true, 64>("tf32 TS 128x64x8   (P V, d = 64)");
  run<true, true, 32>("tf32 TS 128x32x8   (P V, d = 32)");
  run<false, false, 128>("bf16 SS 128x128x16 (S = Q K^T, bf16)");
  run<false, true, 128>("bf16 TS 128x128x16 (P V, d = 128)");
  run<false, true, 64>("bf16 TS 128x64x16  (P V, d = 64)");
  for (int w : {0, 4, 8}) run_contended<true, false, 128>("tf32 SS 128x128x8", w);
  for (int w : {0, 4, 8}) run_contended<true, true, 64>("tf32 TS 128x64x8", w);
  for (int w : {0, 4, 8}) run_contended<false, false, 128>("bf16 SS 128x128x16", w);
  for (int w : {0, 4, 8}) run_contended<false, true, 128>("bf16 TS 128x128x16", w);
  run_pair<true, false, 128>("pair tf32 SS 256x128x8  (S of two CTAs' Q tiles in one MMA)");
  run_pair<true, true, 64>("pair tf32 TS 256x64x8   (P V, d = 64)");
  run_pair<true, true, 32>("pair tf32 TS 256x32x8   (P V, d = 32)");
  run_pair<false, false, 128>("pair bf16 SS 256x128x16");
  run_pair<false, true, 128>("pair bf16 TS 256x128x16 (P V, d = 128)");
  return 0;
}
