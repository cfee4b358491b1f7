// fa_simt.cuh — general-shape CUDA-core forward (any head_dim % 8 == 0 up to 256, any n_q / n_k, strided).
//
// It exists for two reasons: head dims the tcgen05 kernel has no instance for, and as the on-GPU
// cross-check of the tcgen05 kernel at sizes the CPU oracle cannot reach.  It is NOT the product path:
// the five BASELINE configs always dispatch to fa_fwd_sm100_kernel.
//
// One warp per query row, 8 rows per CTA.  K/V tiles of 32 keys are staged in SMEM and shared by the
// 8 warps; lane = key for S = q.k, then the running max / sum are warp-shuffle reductions — the
// "threads collaborate on the row max, log2 instead of linear" item of the reference's own TODO
// (README.md:31; its serial loop is src/flashattention.cu:265-274).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fa {

struct SimtParams {
  const void* q; const void* k; const void* v; void* o; float* lse;
  int64_t q_sb, q_sh, q_sn, k_sb, k_sh, k_sn, v_sb, v_sh, v_sn, o_sb, o_sh, o_sn;  // element strides
  int n_q, n_k, heads, batch, head_dim;
  int kv_group;   // query heads per K/V head (1 = every query head has its own)
  int causal, causal_offset;
  float scale;
};

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float x);
template <> __device__ __forceinline__ void st_from_float<__half>(__half* p, float x) { *p = __float2half_rn(x); }
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float x) { *p = x; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float x) { *p = __float2bfloat16_rn(x); }

constexpr int kSimtRows = 8;     // warps (= query rows) per CTA
constexpr int kSimtKeys = 32;    // keys per staged tile
constexpr int kSimtMaxD = 256;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kSimtRows * 32)
fa_fwd_simt_kernel(const SimtParams p) {
  extern __shared__ float simt_smem[];
  const int d = p.head_dim;
  float* sK = simt_smem;                              // [32][d + 1]
  float* sV = sK + kSimtKeys * (d + 1);               // [32][d]
  float* sQ = sV + kSimtKeys * d;                     // [8][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh % p.heads;
  const int row = blockIdx.x * kSimtRows + warp;
  const bool row_ok = row < p.n_q;
  const TIn* q = static_cast<const TIn*>(p.q) + b * p.q_sb + h * p.q_sh;
  const TIn* k = static_cast<const TIn*>(p.k) + b * p.k_sb + (h / p.kv_group) * p.k_sh;
  const TIn* v = static_cast<const TIn*>(p.v) + b * p.v_sb + (h / p.kv_group) * p.v_sh;

  for (int i = lane; i < d; i += 32) sQ[warp * d + i] = row_ok ? ld_as_float(q + (int64_t)row * p.q_sn + i) : 0.f;

  float m = -INFINITY, l = 0.f;
  float acc[kSimtMaxD / 32];
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) acc[i] = 0.f;

  // keys needed by this CTA (causal: up to the last row of the CTA)
  int k_end = p.n_k;
  if (p.causal) {
    const int last_row = min(blockIdx.x * kSimtRows + kSimtRows - 1, p.n_q - 1);
    k_end = max(0, min(p.n_k, last_row + p.causal_offset + 1));
  }
  const int my_last = p.causal ? row + p.causal_offset : p.n_k - 1;  // last visible key of this row

  for (int k0 = 0; k0 < k_end; k0 += kSimtKeys) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kSimtKeys * d; idx += blockDim.x) {
      const int kr = idx / d, kc = idx % d;
      const int key = k0 + kr;
      float kvv = 0.f, vvv = 0.f;
      if (key < p.n_k) {
        kvv = ld_as_float(k + (int64_t)key * p.k_sn + kc);
        vvv = ld_as_float(v + (int64_t)key * p.v_sn + kc);
      }
      sK[kr * (d + 1) + kc] = kvv;
      sV[kr * d + kc] = vvv;
    }
    __syncthreads();
    const int key = k0 + lane;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s = fmaf(sQ[warp * d + i], sK[lane * (d + 1) + i], s);
    s *= p.scale;
    if (key >= p.n_k || key > my_last) s = -INFINITY;
    // warp-shuffle row max
    float mx = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m_new = fmaxf(m, mx);
    const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = __expf(m - m_safe);          // m = -inf -> 0
    const float pj = __expf(s - m_safe);             // s = -inf -> 0
    float ps = pj;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
    l = l * alpha + ps;
    m = m_new;
#pragma unroll
    for (int i = 0; i < kSimtMaxD / 32; ++i) acc[i] *= alpha;
    for (int kk = 0; kk < kSimtKeys; ++kk) {
      const float pk = __shfl_sync(0xffffffffu, pj, kk);
#pragma unroll
      for (int i = 0; i < kSimtMaxD / 32; ++i) {
        const int col = lane + 32 * i;
        if (col < d) acc[i] = fmaf(pk, sV[kk * d + col], acc[i]);
      }
    }
  }
  if (row_ok) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    TOut* o = static_cast<TOut*>(p.o) + b * p.o_sb + h * p.o_sh + (int64_t)row * p.o_sn;
#pragma unroll
    for (int i = 0; i < kSimtMaxD / 32; ++i) {
      const int col = lane + 32 * i;
      if (col < d) st_from_float(o + col, acc[i] * inv);
    }
    if (p.lse != nullptr && lane == 0)
      p.lse[((int64_t)b * p.heads + h) * p.n_q + row] = l > 0.f ? m + logf(l) : -INFINITY;
  }
}

// ---- log-sum-exp merge of two partials (ring attention), in place on the accumulator ----
__global__ void fa_merge_kernel(float* __restrict__ o_acc, float* __restrict__ lse_acc, const float* __restrict__ o_new,
                                const float* __restrict__ lse_new, int64_t rows, int d4 /* head_dim / 4 */) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = rows * d4;
  if (idx >= total) return;
  const int64_t row = idx / d4;
  const float la = lse_acc[row], lb = lse_new[row];
  const float mx = fmaxf(la, lb);
  float wa, wb;
  if (mx == -INFINITY) {
    wa = 0.f; wb = 0.f;
  } else {
    const float ea = __expf(la - mx), eb = __expf(lb - mx);
    const float inv = 1.f / (ea + eb);
    wa = ea * inv; wb = eb * inv;
  }
  float4 a = reinterpret_cast<float4*>(o_acc)[idx];
  const float4 bvec = reinterpret_cast<const float4*>(o_new)[idx];
  a.x = a.x * wa + bvec.x * wb;
  a.y = a.y * wa + bvec.y * wb;
  a.z = a.z * wa + bvec.z * wb;
  a.w = a.w * wa + bvec.w * wb;
  reinterpret_cast<float4*>(o_acc)[idx] = a;
}
// second pass (after every O element has consumed the old lse): lse_acc = logaddexp(lse_acc, lse_new)
__global__ void fa_merge_lse_kernel(float* __restrict__ lse_acc, const float* __restrict__ lse_new, int64_t rows) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float la = lse_acc[row], lb = lse_new[row];
  const float mx = fmaxf(la, lb);
  lse_acc[row] = (mx == -INFINITY) ? -INFINITY : mx + logf(__expf(la - mx) + __expf(lb - mx));
}

// ---- split-KV across CTAs: merge the kv_splits normalised partials of every query row (workspace slices [s][b, h, i, :] fp32 and
// their LSEs) into the caller's O (its dtype and strides) and LSE.  One warp per row; a run that saw no key has LSE = -inf.
template <typename TOut>
__global__ void __launch_bounds__(256)
fa_combine_splits_kernel(const float* __restrict__ o_ws, const float* __restrict__ lse_ws, int splits, int64_t rows, int heads, int n_q, int d,
                         TOut* __restrict__ o, int64_t o_sb, int64_t o_sh, int64_t o_sn, float* __restrict__ lse) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, lse_ws[s * rows + row]);
  float denom = 0.f;
  if (m != -INFINITY)
    for (int s = 0; s < splits; ++s) denom += __expf(lse_ws[s * rows + row] - m);
  const float inv = denom > 0.f ? 1.f / denom : 0.f;
  float acc[kSimtMaxD / 32];
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) acc[i] = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ls = lse_ws[s * rows + row];
    if (ls == -INFINITY) continue;
    const float w = __expf(ls - m) * inv;
    const float* src = o_ws + (s * rows + row) * d;
#pragma unroll
    for (int i = 0; i < kSimtMaxD / 32; ++i) {
      const int col = lane + 32 * i;
      if (col < d) acc[i] = fmaf(w, src[col], acc[i]);
    }
  }
  const int64_t i_q = row % n_q, bh = row / n_q;
  TOut* dst = o + (bh / heads) * o_sb + (bh % heads) * o_sh + i_q * o_sn;
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) {
    const int col = lane + 32 * i;
    if (col < d) st_from_float(dst + col, acc[i]);
  }
  if (lse != nullptr && lane == 0) lse[row] = denom > 0.f ? m + logf(denom) : -INFINITY;
}

// fp32 -> bf16 / fp16 (final cast of the ring accumulator)
template <typename T16>
__global__ void fa_cast_16_kernel(const float* __restrict__ src, T16* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 x = *reinterpret_cast<const float4*>(src + i);
    alignas(8) T16 y[4];
    st_from_float(&y[0], x.x); st_from_float(&y[1], x.y); st_from_float(&y[2], x.z); st_from_float(&y[3], x.w);
    *reinterpret_cast<uint2*>(dst + i) = *reinterpret_cast<const uint2*>(y);   // one 8-byte store
  } else {
    for (int64_t j = i; j < n; ++j) st_from_float(dst + j, src[j]);
  }
}

}  // namespace fa
