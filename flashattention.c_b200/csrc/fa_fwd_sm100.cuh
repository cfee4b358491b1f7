// fa_fwd_sm100.cuh — FlashAttention forward for sm_100a: TMA -> SMEM ring -> tcgen05.mma -> TMEM.
//
// Replaces the reference's hot loop (src/flashattention.cu:139-355 non-causal, 359-579 causal;
// the llm.c twin src/llm.c/attention_forward.cu:881-1104):
//   tile loader   (FA:217-234, 313-324)  -> one TMA producer thread, mbarrier full/empty ring
//   S = Q K^T     (FA:236-252)           -> tcgen05.mma kind::tf32 / kind::f16, A and B from SMEM, D in TMEM
//   online softmax(FA:258-290)           -> one thread per S row (tcgen05.ld 32x32b), exp2 domain, lazy rescale
//   O += P V      (FA:326-340)           -> tcgen05.mma, A = P read straight from TMEM, B = V (MN-major) from SMEM
//   epilogue      (FA:346-354)           -> O/l -> swizzled SMEM -> TMA store (128-byte coalesced)
//
// One CTA owns 256 query rows of one (batch, head) as two 128-row tiles A and B that ping-pong on the
// tensor pipe: while the softmax warps of one tile work on S_j, the MMA thread runs P*V and the next
// Q*K^T of the other tile.
//
// Warp roles (320 threads): warps 0-3 softmax/epilogue of tile A, 4-7 of tile B (warp w reads TMEM lanes
// 32*(w%4)..+31), warp 8 lane 0 = TMA producer (+ TMEM alloc/dealloc by the whole warp), warp 9 lane 0 = MMA issuer.
//
// TMEM columns (512 allocated): S_A [0,128)  S_B [128,256)  O_A [256,256+d)  O_B [256+d, 256+2d).
// P aliases S: bf16 path packs two bf16 per column into S cols [0,64); tf32 path overwrites S in place.
//
// SMEM (dynamic, 1024-B aligned): Q_A | Q_B | ring of NBUF K/V tiles | barriers.  Every tile is DCHUNKS
// boxes of [128 rows x 128 bytes] in the SWIZZLE_128B layout that TMA writes and the UMMA descriptors read.
#pragma once
#include "ptx.cuh"

namespace fa {

struct FwdParams {
  float scale;       // multiplies q.k
  float scale_log2;  // scale * log2(e)
  int n_q, n_k, heads, batch;
  int causal_offset;  // n_k - n_q
  int num_m_blocks;   // ceil(n_q / 256)
  float* lse;         // [batch, heads, n_q] or nullptr
  uint64_t v_desc_hi; // upper descriptor bits (LBO/SBO/layout) of V as the MN-major B operand of P*V
};

constexpr int kBlockM = 128;          // rows per Q tile
constexpr int kBlockN = 128;          // keys per K/V tile
constexpr int kChunkBytes = 128 * 128;  // one TMA box: 128 rows x 128 bytes
constexpr int kNumThreads = 320;
constexpr float kRescaleThreshold = 8.0f;  // lazy rescale: keep a stale max while it is within 2^8

template <bool kTF32, int kHeadDim, bool kOutF32>
struct FwdTraits {
  static constexpr int kInSize = kTF32 ? 4 : 2;
  static constexpr int kOutSize = (kTF32 || kOutF32) ? 4 : 2;
  static constexpr int kDChunks = kHeadDim * kInSize / 128;    // boxes per Q/K/V tile
  static constexpr int kOChunks = kHeadDim * kOutSize / 128;   // boxes per O tile
  static constexpr int kElemsPerChunk = 128 / kInSize;
  static constexpr int kOutElemsPerChunk = 128 / kOutSize;
  static constexpr int kTileBytes = kDChunks * kChunkBytes;
  static constexpr int kNBuf = kDChunks == 1 ? 8 : 5;          // K/V ring depth (tiles)
  static constexpr int kUmmaK = 32 / kInSize;                  // K per tcgen05.mma: 8 (tf32) / 16 (bf16)
  static constexpr int kSmemData = (2 + kNBuf) * kTileBytes;
  static constexpr int kNumBarriers = 2 /*q*/ + 2 * kNBuf + 2 /*s_full*/ + 2 /*p_full*/ + 2 /*o_final*/;
  static constexpr int kSmemBytes = kSmemData + kNumBarriers * 8 + 16 /*tmem ptr*/ + 1024 /*alignment slack*/;
  static constexpr int kTmemS = 0;        // + 128*t
  static constexpr int kTmemO = 256;      // + kHeadDim*t
  static_assert(kDChunks == 1 || kDChunks == 2, "tile row must be 128 or 256 bytes");
  static_assert(256 + 2 * kHeadDim <= 512, "TMEM budget");
};

// watchdog tags
enum : uint32_t {
  TAG_Q_FULL = 1, TAG_KV_FULL = 2, TAG_KV_EMPTY = 3, TAG_S_FULL = 4, TAG_P_FULL = 5, TAG_O_FINAL = 6
};

template <bool kTF32, int kHeadDim, bool kCausal, bool kOutF32>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                    const FwdParams p) {
  using T = FwdTraits<kTF32, kHeadDim, kOutF32>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                        // 2 tiles
  const uint32_t sKV = smem_base + 2 * T::kTileBytes;   // kNBuf tiles
  const uint32_t sBar = smem_base + T::kSmemData;
  const uint32_t bar_q = sBar;                          // [2]
  const uint32_t bar_full = sBar + 16;                  // [kNBuf]
  const uint32_t bar_empty = bar_full + 8 * T::kNBuf;   // [kNBuf]
  const uint32_t bar_s = bar_empty + 8 * T::kNBuf;      // [2]
  const uint32_t bar_p = bar_s + 16;                    // [2]
  const uint32_t bar_o = bar_p + 16;                    // [2]
  const uint32_t s_tmem_ptr = bar_o + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work assignment: blockIdx.x -> (m block, head, batch); m fastest so neighbours share K/V in L2 ----
  int bid = blockIdx.x;
  int m_blk = bid % p.num_m_blocks;
  bid /= p.num_m_blocks;
  const int head = bid % p.heads;
  const int batch = bid / p.heads;
  if (kCausal) m_blk = p.num_m_blocks - 1 - m_blk;  // heaviest blocks first
  const int row0 = m_blk * (2 * kBlockM);

  // KV trip count per Q tile
  const int n_kv_total = (p.n_k + kBlockN - 1) / kBlockN;
  int n_tile[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = row0 + t * kBlockM;
    int n = n_kv_total;
    if (kCausal) {
      const int last_key = r0 + kBlockM - 1 + p.causal_offset;
      n = last_key < 0 ? 0 : min(n_kv_total, last_key / kBlockN + 1);
    }
    if (r0 >= p.n_q) n = 0;
    n_tile[t] = n;
  }
  const int n_max = max(n_tile[0], n_tile[1]);

  // ---- one-time setup ----
  if (warp == 9 && lane == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_q + 8, 1);
    for (int i = 0; i < T::kNBuf; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_s + 8 * t, 1);
      mbar_init(bar_p + 8 * t, 128);
      mbar_init(bar_o + 8 * t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
      tma_prefetch_desc(&tm_o);
    }
    tmem_alloc(s_tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem_ptr));

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0 && n_max > 0) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (n_tile[t] > 0) {
          mbar_arrive_expect_tx(bar_q + 8 * t, T::kTileBytes);
#pragma unroll
          for (int c = 0; c < T::kDChunks; ++c)
            tma_load_4d(sQ + t * T::kTileBytes + c * kChunkBytes, &tm_q, bar_q + 8 * t, c * T::kElemsPerChunk,
                        row0 + t * kBlockM, head, batch);
        }
      }
      const int n_loads = 2 * n_max;  // K_0, V_0, K_1, V_1, ...
      for (int i = 0; i < n_loads; ++i) {
        const int buf = i % T::kNBuf;
        const int round = i / T::kNBuf;
        if (round > 0) mbar_wait(bar_empty + 8 * buf, (round - 1) & 1, TAG_KV_EMPTY);
        mbar_arrive_expect_tx(bar_full + 8 * buf, T::kTileBytes);
        const CUtensorMap* tm = (i & 1) ? &tm_v : &tm_k;
        const int kv0 = (i >> 1) * kBlockN;
#pragma unroll
        for (int c = 0; c < T::kDChunks; ++c)
          tma_load_4d(sKV + buf * T::kTileBytes + c * kChunkBytes, tm, bar_full + 8 * buf, c * T::kElemsPerChunk, kv0,
                      head, batch);
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && n_max > 0) {
      constexpr uint32_t kFmt = kTF32 ? 2u : 1u;
      constexpr uint32_t idesc_s = make_idesc(kFmt, 0, kBlockM, kBlockN);
      constexpr uint32_t idesc_pv = make_idesc(kFmt, 1, kBlockM, kHeadDim);
      // K-major operands (Q, K): LBO unused for swizzled K-major (encoded 1), SBO = 1024 B between 8-row groups
      constexpr uint64_t hi_kmajor = make_sdesc_hi_sw128(16, 1024);
      // MN-major operand (V as B of P*V): LBO = stride between 128-byte column chunks (one TMA box), SBO = stride
      // between key groups.  bf16: SWIZZLE_128B, 8-key groups of 1024 B.  tf32: tcgen05 only accepts the
      // SWIZZLE_128B_BASE32B layout for MN-major 32-bit operands (4-key groups of 512 B), which TMA writes with
      // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Built on the host next to the matching tensor map.
      const uint64_t hi_mnmajor = p.v_desc_hi;
      constexpr int kKStepsS = kHeadDim / T::kUmmaK;   // k-steps of Q K^T (32 bytes each)
      constexpr int kKStepsPV = kBlockN / T::kUmmaK;   // k-steps of P V (UmmaK keys each)

      auto issue_s = [&](int t, int buf) {
        const uint32_t qa = sQ + t * T::kTileBytes;
        const uint32_t kb = sKV + buf * T::kTileBytes;
        const uint32_t d = tmem_base + T::kTmemS + t * kBlockN;
#pragma unroll
        for (int kk = 0; kk < kKStepsS; ++kk) {
          const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
          mma_ss<kTF32>(d, sdesc_at(hi_kmajor, qa + off), sdesc_at(hi_kmajor, kb + off), idesc_s, kk > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int buf, bool accumulate) {
        const uint32_t vb = sKV + buf * T::kTileBytes;
        const uint32_t d = tmem_base + T::kTmemO + t * kHeadDim;
        const uint32_t a = tmem_base + T::kTmemS + t * kBlockN;  // P aliases S
#pragma unroll
        for (int ks = 0; ks < kKStepsPV; ++ks) {
          mma_ts<kTF32>(d, a + ks * 8, sdesc_at(hi_mnmajor, vb + ks * (T::kUmmaK * 128)), idesc_pv,
                        (accumulate || ks > 0) ? 1u : 0u);
        }
      };

      if (n_tile[0] > 0) mbar_wait(bar_q, 0, TAG_Q_FULL);
      if (n_tile[1] > 0) mbar_wait(bar_q + 8, 0, TAG_Q_FULL);
      // prologue: S_t(0) = Q_t K_0^T
      mbar_wait(bar_full + 0, 0, TAG_KV_FULL);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (n_tile[t] > 0) {
          issue_s(t, 0);
          tc_commit(bar_s + 8 * t);
        }
      }
      tc_commit(bar_empty + 0);

      for (int j = 0; j < n_max; ++j) {
        const int iv = 2 * j + 1;   // ring index of V_j
        const int ik = 2 * j + 2;   // ring index of K_{j+1}
        const int vbuf = iv % T::kNBuf;
        const int kbuf = ik % T::kNBuf;
        mbar_wait(bar_full + 8 * vbuf, (iv / T::kNBuf) & 1, TAG_KV_FULL);
        bool k_ready = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j < n_tile[t]) {
            mbar_wait(bar_p + 8 * t, j & 1, TAG_P_FULL);
            tc_fence_after();
            issue_pv(t, vbuf, j > 0);
            if (j == n_tile[t] - 1) tc_commit(bar_o + 8 * t);
            if (j + 1 < n_tile[t]) {
              if (!k_ready) {
                mbar_wait(bar_full + 8 * kbuf, (ik / T::kNBuf) & 1, TAG_KV_FULL);
                tc_fence_after();
                k_ready = true;
              }
              issue_s(t, kbuf);
              tc_commit(bar_s + 8 * t);
            }
          }
        }
        tc_commit(bar_empty + 8 * vbuf);
        if (j + 1 < n_max) tc_commit(bar_empty + 8 * kbuf);
      }
    }
  } else {
    // =========================== softmax + epilogue (warps 0-7) ===========================
    const int t = warp >> 2;                       // Q tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;          // row within the tile == TMEM lane
    const int q_row = row0 + t * kBlockM + r;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + T::kTmemS + t * kBlockN;
    const uint32_t tO = tmem_base + lane_base + T::kTmemO + t * kHeadDim;
    const int n_mine = n_tile[t];
    const float c = p.scale_log2;

    float m = -INFINITY;  // running (possibly stale) row max, in raw q.k units
    float l = 0.f;        // running row sum of exp2((s - m) * c)

    for (int j = 0; j < n_mine; ++j) {
      mbar_wait(bar_s + 8 * t, j & 1, TAG_S_FULL);
      tc_fence_after();
      float s[128];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld32(tS + q4 * 32, reinterpret_cast<uint32_t*>(&s[q4 * 32]));
      tc_wait_ld();

      // masking: key kv0 + i is visible iff i <= limit
      const int kv0 = j * kBlockN;
      int limit = p.n_k - 1 - kv0;
      if (kCausal) limit = min(limit, q_row + p.causal_offset - kv0);
      if (limit < kBlockN - 1) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i > limit) s[i] = -INFINITY;
      }

      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        mx0 = fmaxf(mx0, s[i]);
        mx1 = fmaxf(mx1, s[i + 1]);
        mx2 = fmaxf(mx2, s[i + 2]);
        mx3 = fmaxf(mx3, s[i + 3]);
      }
      const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));

      if (j == 0) {
        m = m_new;
      } else {
        // lazy rescale: only move the reference max when it grew by more than 2^kRescaleThreshold
        const bool need = (m_new - m) * c > kRescaleThreshold;   // (-inf -> finite) gives +inf -> true
        if (__any_sync(0xffffffffu, need)) {
          const float m_use = need ? m_new : m;
          const float alpha = need ? ex2((m - m_use) * c) : 1.0f;  // m = -inf -> 0
          l *= alpha;
#pragma unroll
          for (int cc = 0; cc < kHeadDim / 16; ++cc) {
            uint32_t o[16];
            tmem_ld16(tO + cc * 16, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + cc * 16, o);
          }
          m = m_use;
        }
      }
      const float m_safe = (m == -INFINITY) ? 0.f : m;
      const float neg_mc = -m_safe * c;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        s[i] = ex2(fmaf(s[i], c, neg_mc));
        s[i + 1] = ex2(fmaf(s[i + 1], c, neg_mc));
        s[i + 2] = ex2(fmaf(s[i + 2], c, neg_mc));
        s[i + 3] = ex2(fmaf(s[i + 3], c, neg_mc));
        if constexpr (kTF32) {
          // kind::tf32 reads only the top 19 bits of P; sum exactly those values so that O = (sum P~ V) / (sum P~)
          // is normalised by what the tensor core actually multiplied (removes the truncation bias from O)
          s[i] = __uint_as_float(__float_as_uint(s[i]) & 0xFFFFE000u);
          s[i + 1] = __uint_as_float(__float_as_uint(s[i + 1]) & 0xFFFFE000u);
          s[i + 2] = __uint_as_float(__float_as_uint(s[i + 2]) & 0xFFFFE000u);
          s[i + 3] = __uint_as_float(__float_as_uint(s[i + 3]) & 0xFFFFE000u);
        }
        l0 += s[i];
        l1 += s[i + 1];
        l2 += s[i + 2];
        l3 += s[i + 3];
      }
      l += (l0 + l1) + (l2 + l3);

      if constexpr (kTF32) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) tmem_st32(tS + q4 * 32, reinterpret_cast<uint32_t*>(&s[q4 * 32]));
      } else {
        uint32_t pk[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) pk[i] = pack_bf16x2(s[2 * i], s[2 * i + 1]);
        tmem_st32(tS, &pk[0]);
        tmem_st32(tS + 32, &pk[32]);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p + 8 * t);
    }

    // ---- epilogue: O/l -> swizzled SMEM (reusing this tile's Q buffer) -> TMA store; LSE -> global ----
    if (n_mine > 0) {
      mbar_wait(bar_o + 8 * t, 0, TAG_O_FINAL);
      tc_fence_after();
    }
    const float inv_l = (n_mine > 0 && l > 0.f) ? 1.0f / l : 0.f;
    if (p.lse != nullptr && q_row < p.n_q) {
      const float m_safe = (m == -INFINITY) ? 0.f : m;
      const float lse = (n_mine > 0 && l > 0.f) ? m_safe * p.scale + logf(l) : -INFINITY;
      p.lse[(static_cast<int64_t>(batch) * p.heads + head) * p.n_q + q_row] = lse;
    }
    const uint32_t stage = sQ + t * T::kTileBytes;       // kDChunks boxes of 16 KB
    const uint32_t row_off = r * 128;
    const uint32_t sw = r & 7;
    constexpr int kRounds = T::kOChunks / T::kDChunks;   // 1, or 2 for bf16-in / fp32-out
    constexpr int kColsPerChunk = T::kOutElemsPerChunk;  // 32 (fp32) or 64 (bf16)
#pragma unroll
    for (int round = 0; round < kRounds; ++round) {
#pragma unroll
      for (int ch = 0; ch < T::kDChunks; ++ch) {
        const int col0 = (round * T::kDChunks + ch) * kColsPerChunk;
#pragma unroll
        for (int half = 0; half < kColsPerChunk / 32; ++half) {
          uint32_t o[32];
          if (n_mine > 0) {
            tmem_ld32(tO + col0 + half * 32, o);
            tc_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * inv_l);
          const uint32_t base = stage + ch * kChunkBytes + row_off;
          if constexpr (T::kOutSize == 4) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              st_shared_v4(base + ((static_cast<uint32_t>(g) ^ sw) << 4), o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t chunk16 = static_cast<uint32_t>(half * 4 + g);
              st_shared_v4(base + ((chunk16 ^ sw) << 4),
                           pack_bf16x2(__uint_as_float(o[8 * g]), __uint_as_float(o[8 * g + 1])),
                           pack_bf16x2(__uint_as_float(o[8 * g + 2]), __uint_as_float(o[8 * g + 3])),
                           pack_bf16x2(__uint_as_float(o[8 * g + 4]), __uint_as_float(o[8 * g + 5])),
                           pack_bf16x2(__uint_as_float(o[8 * g + 6]), __uint_as_float(o[8 * g + 7])));
            }
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if ((warp & 3) == 0 && lane == 0) {
#pragma unroll
        for (int ch = 0; ch < T::kDChunks; ++ch)
          tma_store_4d(&tm_o, stage + ch * kChunkBytes, (round * T::kDChunks + ch) * kColsPerChunk, row0 + t * kBlockM, head,
                       batch);
        tma_store_commit();
        tma_store_wait_read();
      }
      if (round + 1 < kRounds) named_bar_sync(1 + t, 128);
    }
    if ((warp & 3) == 0 && lane == 0) tma_store_wait_all();
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fa
