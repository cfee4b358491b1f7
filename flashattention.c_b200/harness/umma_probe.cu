// umma_probe.cu — single-CTA known-answer tests of the tcgen05 building blocks the forward kernel is
// assembled from, with the descriptor fields supplied at RUN time so one GPU session can try several
// hypotheses:   ./umma_probe <dtype: tf32|bf16> <d> <mode: ss|ts> <variant>
//   ss: S[128x128]  = Q[128 x d] * K[128 x d]^T     both operands K-major SWIZZLE_128B from TMA
//   ts: O[128 x d]  = P[128 x 128] (TMEM) * V[128 x d]   V MN-major from TMA
//   sv: O[128 x d]  = P[128 x 128] (SMEM, K-major) * V[128 x d]   isolates the V descriptor from the TMEM-A layout
// Inputs are small integers (exact in tf32 and bf16), so the expected error is exactly 0.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../csrc/ptx.cuh"

using namespace fa;

struct ProbeParams {
  int n_ksteps;
  int n_out_cols;       // columns of D to read back
  int ts_mode;          // 0: A from smem, 1: A (=P) from TMEM
  int pack_swap;        // bf16 TS: swap lo/hi halves when packing P
  uint32_t idesc;
  uint64_t a_hi, b_hi;  // smem descriptor templates
  uint32_t a_off[16];   // per k-step: smem byte offset (ss) or TMEM column offset (ts)
  uint32_t b_off[16];
  int a_bytes, b_bytes; // TMA bytes for A and B
  int a_chunks, b_chunks, elems_per_chunk;
};

template <bool kTF32>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ProbeParams pp,
             const float* __restrict__ p_in, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 65536, bar_ld = base + 131072, bar_mma = bar_ld + 8, s_tptr = bar_ld + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(s_tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tptr));
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_ld, pp.a_bytes + pp.b_bytes);
    for (int c = 0; c < pp.a_chunks; ++c) tma_load_4d(sA + c * 16384, &tmA, bar_ld, c * pp.elems_per_chunk, 0, 0, 0);
    for (int c = 0; c < pp.b_chunks; ++c) tma_load_4d(sB + c * 16384, &tmB, bar_ld, c * pp.elems_per_chunk, 0, 0, 0);
  }
  if (pp.ts_mode) {
    // every thread writes its row of P into TMEM columns [0, 128) (tf32) or [0, 64) (bf16 pairs)
    const float* prow = p_in + (size_t)threadIdx.x * 128;
    if constexpr (kTF32) {
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(prow[q * 32 + i]);
        tmem_st32(tmem_base + lane_base + q * 32, r);
      }
    } else {
      for (int q = 0; q < 2; ++q) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float a = prow[q * 64 + 2 * i], b = prow[q * 64 + 2 * i + 1];
          r[i] = pp.pack_swap ? pack_bf16x2(b, a) : pack_bf16x2(a, b);
        }
        tmem_st32(tmem_base + lane_base + q * 32, r);
      }
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_wait(bar_ld, 0, 100);
    tc_fence_after();
    const uint32_t d = tmem_base + 256;
    for (int ks = 0; ks < pp.n_ksteps; ++ks) {
      if (pp.ts_mode)
        mma_ts<kTF32>(d, tmem_base + pp.a_off[ks], sdesc_at(pp.b_hi, sB + pp.b_off[ks]), pp.idesc, ks > 0);
      else
        mma_ss<kTF32>(d, sdesc_at(pp.a_hi, sA + pp.a_off[ks]), sdesc_at(pp.b_hi, sB + pp.b_off[ks]), pp.idesc, ks > 0);
    }
    tc_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0, 101);
  tc_fence_after();
  for (int c0 = 0; c0 < pp.n_out_cols; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + lane_base + 256 + c0, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(size_t)threadIdx.x * pp.n_out_cols + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);  // inputs are exactly representable
}

int main(int argc, char** argv) {
  if (argc < 5) {
    printf("usage: %s tf32|bf16 d ss|ts variant\n", argv[0]);
    return 1;
  }
  const bool tf32 = !strcmp(argv[1], "tf32");
  const int d = atoi(argv[2]);
  const bool ts = !strcmp(argv[3], "ts");
  const bool sv = !strcmp(argv[3], "sv");
  const bool pv = ts || sv;
  const int variant = atoi(argv[4]);
  const int es = tf32 ? 4 : 2;
  const int epc = 128 / es;            // elements per 128-byte chunk row
  const int chunks = d * es / 128;     // boxes per tile
  const int umma_k = 32 / es;

  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;

  // host data: A = Q or unused (ts), B = K (ss) or V (ts), both [128 rows x d]
  std::vector<float> A(128 * d), B(128 * d), P(128 * 128);
  uint32_t seed = 12345u;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (int)((seed >> 24) % 9) - 4; };
  for (auto& x : A) x = (float)rnd();
  for (auto& x : B) x = (float)rnd();
  for (auto& x : P) x = (float)(rnd() + 4) * 0.125f;
  std::vector<uint8_t> Ab(128 * d * es), Bb(128 * d * es);
  for (int i = 0; i < 128 * d; ++i) {
    if (tf32) { memcpy(&Ab[i * 4], &A[i], 4); memcpy(&Bb[i * 4], &B[i], 4); }
    else { uint16_t a = f2bf(A[i]), b = f2bf(B[i]); memcpy(&Ab[i * 2], &a, 2); memcpy(&Bb[i * 2], &b, 2); }
  }
  void *dA, *dB; float *dP, *dOut;
  const int n_out = pv ? d : 128;
  if (sv) {  // A operand = P [128 x 128] in the input dtype, K-major
    Ab.assign(128 * 128 * es, 0);
    for (int i = 0; i < 128 * 128; ++i) {
      if (tf32) memcpy(&Ab[i * 4], &P[i], 4);
      else { uint16_t a = f2bf(P[i]); memcpy(&Ab[i * 2], &a, 2); }
    }
  }
  CK(cudaMalloc(&dB, Bb.size())); CK(cudaMalloc(&dP, P.size() * 4));
  CK(cudaMalloc(&dOut, 128 * n_out * 4));
  CK(cudaMemcpy(dB, Bb.data(), Bb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0xff, 128 * n_out * 4));

  auto mk = [&](CUtensorMap* m, void* ptr, int inner, CUtensorMapSwizzle swz) {
    cuuint64_t dims[4] = {(cuuint64_t)inner, 128, 1, 1};
    cuuint64_t str[3] = {(cuuint64_t)inner * es, (cuuint64_t)128 * inner * es, (cuuint64_t)128 * inner * es};
    cuuint32_t box[4] = {(cuuint32_t)epc, 128, 1, 1}, estr[4] = {1, 1, 1, 1};
    return enc(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, str, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  ProbeParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.ts_mode = ts; pp.n_out_cols = n_out; pp.a_chunks = sv ? 128 * es / 128 : chunks; pp.b_chunks = chunks; pp.elems_per_chunk = epc;
  pp.a_bytes = pp.a_chunks * 16384; pp.b_bytes = chunks * 16384;
  const uint32_t fmt = tf32 ? 2u : 1u;
  const char* vname = "";
  CUtensorMapSwizzle b_swz = CU_TENSOR_MAP_SWIZZLE_128B;
  if (!pv) {
    pp.idesc = make_idesc(fmt, 0, 128, 128);
    pp.n_ksteps = d / umma_k;
    for (int k = 0; k < pp.n_ksteps; ++k) pp.a_off[k] = pp.b_off[k] = (k >> 2) * 16384 + (k & 3) * 32;
    uint32_t lbo = 16, sbo = 1024;
    switch (variant) {
      case 0: vname = "K-major lbo=16 sbo=1024 (expected)"; break;
      case 1: vname = "K-major lbo=0 sbo=1024"; lbo = 0; break;
      default: printf("bad variant\n"); return 1;
    }
    pp.a_hi = pp.b_hi = make_sdesc_hi_sw128(lbo, sbo);
  } else {
    pp.idesc = make_idesc(fmt, 1, 128, d);
    pp.n_ksteps = 128 / umma_k;
    uint32_t lbo = 16384, sbo = 1024, lt = kLayoutSw128;
    switch (variant) {
      case 0: vname = "V MN-major SW128 lbo=16384 sbo=1024"; break;
      case 1: vname = "V MN-major SW128_BASE32B (TMA ATOM_32B) lbo=16384 sbo=512"; lt = kLayoutSw128Base32; sbo = 512; b_swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; break;
      case 2: vname = "V MN-major SW128_BASE32B (TMA ATOM_32B) lbo=16384 sbo=1024"; lt = kLayoutSw128Base32; b_swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; break;
      case 3: vname = "V MN-major SW128_BASE32B (TMA ATOM_32B) lbo=512 sbo=16384"; lt = kLayoutSw128Base32; lbo = 512; sbo = 16384; b_swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; break;
      case 4: vname = "V MN-major SW128_BASE32B desc over TMA SWIZZLE_128B data lbo=16384 sbo=512"; lt = kLayoutSw128Base32; sbo = 512; break;
      default: printf("bad variant\n"); return 1;
    }
    pp.b_hi = make_sdesc_hi(lbo, sbo, lt);
    pp.a_hi = make_sdesc_hi_sw128(16, 1024);
    for (int k = 0; k < pp.n_ksteps; ++k) {
      pp.a_off[k] = ts ? k * 8 : (k >> 2) * 16384 + (k & 3) * 32;
      pp.b_off[k] = k * umma_k * 128;
    }
  }
  CUtensorMap tmA, tmB;
  CK(cudaMalloc(&dA, Ab.size()));
  CK(cudaMemcpy(dA, Ab.data(), Ab.size(), cudaMemcpyHostToDevice));
  if (mk(&tmA, dA, sv ? 128 : d, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS || mk(&tmB, dB, d, b_swz) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 2;
  }

  const int smem = 131072 + 64 + 1024;
  if (tf32) {
    CK(cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<true><<<1, 128, smem>>>(tmA, tmB, pp, dP, dOut);
  } else {
    CK(cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<false><<<1, 128, smem>>>(tmA, tmB, pp, dP, dOut);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PROBE %s d=%d %s v%d [%s]: CUDA ERROR %s\n", argv[1], d, argv[3], variant, vname, cudaGetErrorString(e));
    return 3;
  }
  std::vector<float> out(128 * n_out);
  CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0; int bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < n_out; ++j) {
      double ref = 0;
      if (!pv) for (int k = 0; k < d; ++k) ref += (double)A[i * d + k] * B[j * d + k];
      else for (int k = 0; k < 128; ++k) ref += (double)P[i * 128 + k] * B[k * d + j];
      const double err = fabs(ref - out[i * n_out + j]);
      if (!(err <= 1e-3)) {
        if (bad < 4) printf("   mismatch [%d][%d] got %f want %f\n", i, j, out[i * n_out + j], ref);
        ++bad;
      }
      if (err > max_err || err != err) max_err = err;
    }
  printf("PROBE %s d=%d %s v%d [%s]: %s max_err=%g bad=%d/%d\n", argv[1], d, argv[3], variant, vname, bad ? "FAIL" : "PASS",
         max_err, bad, 128 * n_out);
  return bad ? 4 : 0;
}
