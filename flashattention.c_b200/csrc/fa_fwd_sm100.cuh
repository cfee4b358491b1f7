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
// 32*(w%4)..+31), warp 8 lane 0 = TMA producer (+ TMEM alloc/dealloc by the whole warp), warp 9 = MMA issuer
// (converged warp, one elected lane issues A then B in order).  `tcgen05.mma` issue back-pressures at the rate the
// tensor pipe retires (~64 cycles per 128x128x16 MMA) and every mbarrier wait + tcgen05 fence costs ~200 cycles
// even when already complete (measured: profiles/r01_trace_v1_c4_report.txt), so waits are batched.  Two
// independent issuer warps (one per tile) were measured and rejected: the tiles fall into lock-step
// (profiles/r01_ab_issuer_modes_session9.log).
//
// Tail CTAs.  The host sizes the grid as whole waves of 256-row CTAs plus a remainder wave of 128-row CTAs.  A
// remainder CTA runs its one Q tile in "split-KV" mode: tile slot A attends the first half of the K/V tiles and
// slot B the second half (same Q rows), so the two halves ping-pong on the tensor pipe exactly like two Q tiles
// do, and the partial (O, m, l) pairs are merged by the log-sum-exp rule through SMEM in the epilogue.  The serial
// chain of a tail CTA is therefore half as long (C1 is nothing but tail CTAs).
//
// TMEM columns (512 allocated): S_A [0,128)  S_B [128,256)  O_A [256,256+d)  O_B [256+d, 256+2d).
// P aliases S: bf16 path packs two bf16 per column into S cols [0,64); tf32 path overwrites S in place.
//
// SMEM (dynamic, 1024-B aligned): Q_A | Q_B | ring of NBUF K/V tiles | barriers.  Every tile is DCHUNKS
// boxes of [128 rows x 128 bytes] in the SWIZZLE_128B layout that TMA writes and the UMMA descriptors read.
#pragma once
#include "ptx.cuh"

// bring-up / tuning switches (A/B-tested on hardware; the defaults are what ships)
#ifndef FA_OPT_SPLITP
#define FA_OPT_SPLITP 1   // deliver P to the MMA thread in two 64-key halves so P*V overlaps the second half of the exps
#endif
#ifndef FA_OPT_LDPIPE
#define FA_OPT_LDPIPE 0   // overlap the row-max pass with the remaining tcgen05.ld of the S row
#endif
#ifndef FA_OPT_F2
#define FA_OPT_F2 1       // packed FFMA2 / FADD2 for the scale-subtract and the row sum (bf16 instances only: in the tf32
                          // instances the per-element P truncation breaks register pairing and costs ~150 extra moves)
#endif
#ifndef FA_OPT_POLY
#define FA_OPT_POLY 0     // of every 8 P elements, how many get exp2 from the FMA-pipe polynomial instead of MUFU.EX2 (0, 2, 4)
#endif
// -DFA_TRACE=1 builds a timeline-tracing kernel: CTA 0 records clock64() at every pipeline hand-off of its first
// kTraceSteps KV tiles into FwdParams::trace ([role 0..3][step][slot 0..7]); see scripts/trace_report.py.
#ifndef FA_TRACE
#define FA_TRACE 0
#endif
#if FA_TRACE
#define FA_TRACE_AT(role, step, slot)                                                                      \
  do {                                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && (step) < fa::kTraceSteps && ((role) < 2 || (threadIdx.x & 31) == 0)) \
      p.trace[((role) * fa::kTraceSteps + (step)) * 8 + (slot)] = static_cast<unsigned long long>(clock64()); \
  } while (0)
#else
#define FA_TRACE_AT(role, step, slot) do { } while (0)
#endif
// one-off CTA milestones go into the last step row of a role: (role 0/1 = softmax A/B: 0 entry, 1 setup done, 2 K/V loop
// done, 3 O final, 4 O staged in SMEM, 5 TMA store issued+read, 6 store complete; role 2 = MMA warp: 0 entry, 1 Q full,
// 2 K_0 full, 3 S_0 issued)
#define FA_TRACE_MISC(role, slot) FA_TRACE_AT(role, fa::kTraceSteps - 1, slot)

namespace fa {

struct FwdParams {
  float scale;       // multiplies q.k
  float scale_log2;  // scale * log2(e)
  int n_q, n_k, heads, batch;
  int causal_offset;  // n_k - n_q
  int num_m_blocks;   // ceil(n_q / 256)
  float* lse;         // [batch, heads, n_q] or nullptr
  uint64_t v_desc_hi; // upper descriptor bits (LBO/SBO/layout) of V as the MN-major B operand of P*V
  unsigned long long* trace;  // FA_TRACE builds only; nullptr otherwise
  int n_big;          // CTAs [0, n_big) own a 256-row block (tiles A+B); CTAs beyond own a 128-row half block
  int tail_split;     // != 0: a 128-row CTA splits its K/V range over tile slots A and B (merged in the epilogue);
                      // 0: it runs slot A only
};
constexpr int kTraceSteps = 48;

constexpr int kBlockM = 128;          // rows per Q tile
constexpr int kBlockN = 128;          // keys per K/V tile
constexpr int kChunkBytes = 128 * 128;  // one TMA box: 128 rows x 128 bytes
constexpr int kNumThreads = 320;      // 8 softmax warps + TMA producer warp + MMA-issuer warp
constexpr int kBarMerge = 3;          // named barrier: slot B hands its partial (O, m, l) to slot A (split-KV tail CTAs)
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
  static constexpr int kNumBarriers = 2 /*q*/ + 2 * kNBuf + 2 /*s_full*/ + 4 /*p_full halves*/ + 2 /*o_final*/;
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
  const uint32_t bar_p = bar_s + 16;                    // [tile][half] = [4]
  const uint32_t bar_o = bar_p + 32;                    // [2]
  const uint32_t s_tmem_ptr = bar_o + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FA_TRACE_MISC(0, 0);
  if (threadIdx.x == 128) FA_TRACE_MISC(1, 0);
  if (warp == 9) FA_TRACE_MISC(2, 0);

  // ---- work assignment: blockIdx.x -> (m block, head, batch); m fastest so neighbours share K/V in L2 ----
  // Wave quantisation: the host sizes n_big to whole waves of 256-row blocks; the remainder blocks (if they are few
  // enough) are issued as pairs of 128-row CTAs so the last, partial wave is short.
  int bid = blockIdx.x;
  const bool single_tile = bid >= p.n_big;
  int half = 0;
  if (single_tile) {
    const int k = bid - p.n_big;
    half = k & 1;
    bid = p.n_big + (k >> 1);
  }
  int m_blk = bid % p.num_m_blocks;
  bid /= p.num_m_blocks;
  const int head = bid % p.heads;
  const int batch = bid / p.heads;
  if (kCausal) m_blk = p.num_m_blocks - 1 - m_blk;  // heaviest blocks first
  const int row0 = m_blk * (2 * kBlockM) + half * kBlockM;
  const bool split = single_tile && p.tail_split != 0;   // slots A and B = two halves of the K/V range of ONE Q tile

  // KV trip count and first K/V tile of each tile slot
  const int n_kv_total = (p.n_k + kBlockN - 1) / kBlockN;
  int n_tile[2];
  int kv_first[2] = {0, 0};
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = row0 + t * kBlockM;
    int n = n_kv_total;
    if (kCausal) {
      const int last_key = r0 + kBlockM - 1 + p.causal_offset;
      n = last_key < 0 ? 0 : min(n_kv_total, last_key / kBlockN + 1);
    }
    if (r0 >= p.n_q || (single_tile && t == 1)) n = 0;
    n_tile[t] = n;
  }
  if (split) {
    const int n = n_tile[0];
    n_tile[0] = (n + 1) >> 1;
    n_tile[1] = n - n_tile[0];
    kv_first[1] = n_tile[0];
  }
  const int n_max = max(n_tile[0], n_tile[1]);
  const int row_of_tile1 = split ? row0 : row0 + kBlockM;   // first Q row of slot B

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
      mbar_init(bar_p + 16 * t, 128);
      mbar_init(bar_p + 16 * t + 8, 128);
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
  if (threadIdx.x == 0) FA_TRACE_MISC(0, 1);
  if (threadIdx.x == 128) FA_TRACE_MISC(1, 1);

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0 && n_max > 0) {
      // Q: one tile per slot; in split mode both slots read the same Q tile from buffer 0
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (n_tile[t] > 0 && !(split && t == 1)) {
          mbar_arrive_expect_tx(bar_q + 8 * t, T::kTileBytes);
#pragma unroll
          for (int c = 0; c < T::kDChunks; ++c)
            tma_load_4d(sQ + t * T::kTileBytes + c * kChunkBytes, &tm_q, bar_q + 8 * t, c * T::kElemsPerChunk,
                        row0 + t * kBlockM, head, batch);
        }
      }
      int ring = 0;   // running index into the K/V ring; the MMA issuer consumes tiles in exactly this order
      auto load = [&](const CUtensorMap* tm, int kv_tile) {
        const int buf = ring % T::kNBuf;
        const int round = ring / T::kNBuf;
        ++ring;
        if (round > 0) mbar_wait(bar_empty + 8 * buf, (round - 1) & 1, TAG_KV_EMPTY);
        mbar_arrive_expect_tx(bar_full + 8 * buf, T::kTileBytes);
#pragma unroll
        for (int c = 0; c < T::kDChunks; ++c)
          tma_load_4d(sKV + buf * T::kTileBytes + c * kChunkBytes, tm, bar_full + 8 * buf, c * T::kElemsPerChunk,
                      kv_tile * kBlockN, head, batch);
      };
      if (!split) {
        for (int j = 0; j < n_max; ++j) {   // K_0, V_0, K_1, V_1, ... shared by both Q tiles
          load(&tm_k, j);
          load(&tm_v, j);
        }
      } else {
        // K_A0, K_B0, then per step: V_A(j), K_A(j+1), V_B(j), K_B(j+1)
        const int nA = n_tile[0], nB = n_tile[1];
        load(&tm_k, 0);
        if (nB > 0) load(&tm_k, nA);
        for (int j = 0; j < nA; ++j) {
          load(&tm_v, j);
          if (j + 1 < nA) load(&tm_k, j + 1);
          if (j < nB) {
            load(&tm_v, nA + j);
            if (j + 1 < nB) load(&tm_k, nA + j + 1);
          }
        }
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    // The whole warp follows the (warp-uniform) control flow and waits on the mbarriers; one elected lane issues.
    if (n_max > 0) {
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
      // Descriptors are built once per operand tile; a k-step only adds to the 14-bit start-address field
      // ((bytes >> 4); SMEM addresses are < 2^18, so the field never carries).
      auto issue_s = [&](int t, int qbuf, int buf) {
        const uint64_t qd = sdesc_at(hi_kmajor, sQ + qbuf * T::kTileBytes);
        const uint64_t kd = sdesc_at(hi_kmajor, sKV + buf * T::kTileBytes);
        const uint32_t d = tmem_base + T::kTmemS + t * kBlockN;
#pragma unroll
        for (int kk = 0; kk < kKStepsS; ++kk) {
          const uint32_t off16 = ((kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4;
          mma_ss<kTF32>(d, qd + off16, kd + off16, idesc_s, kk > 0 ? 1u : 0u);
        }
      };
      // P*V for k-steps [ks0, ks1) of the 128-key tile; A = P_t read from TMEM (it aliases S_t)
      auto issue_pv = [&](int t, int buf, bool accumulate, int ks0, int ks1) {
        const uint64_t vd = sdesc_at(hi_mnmajor, sKV + buf * T::kTileBytes);
        const uint32_t d = tmem_base + T::kTmemO + t * kHeadDim;
        const uint32_t a = tmem_base + T::kTmemS + t * kBlockN;
#pragma unroll
        for (int ks = ks0; ks < ks1; ++ks) {
          mma_ts<kTF32>(d, a + ks * 8, vd + static_cast<uint32_t>(ks * (T::kUmmaK * 128 / 16)), idesc_pv,
                        (accumulate || ks > 0) ? 1u : 0u);
        }
      };
      auto wait_full = [&](int i) { mbar_wait(bar_full + 8 * (i % T::kNBuf), (i / T::kNBuf) & 1, TAG_KV_FULL); };
      // P_t(j) V -> O_t in two 64-key halves as the softmax warps deliver them, then (unless this was the slot's last
      // K/V tile) S_t(j+1); the tensor pipe executes in issue order, so S_t(j+1) may overwrite the columns P_t(j) aliased.
      auto step = [&](int t, int j, bool last, int qbuf, int vbuf, int kbuf, bool release) {
#if FA_OPT_SPLITP
        FA_TRACE_AT(2 + t, j, 0);
        mbar_wait(bar_p + 16 * t, j & 1, TAG_P_FULL);          // keys [0, 64) of P are in TMEM
        tc_fence_after();
        FA_TRACE_AT(2 + t, j, 1);
        if (elect_one_sync()) issue_pv(t, vbuf, j > 0, 0, kKStepsPV / 2);
        __syncwarp();
        FA_TRACE_AT(2 + t, j, 2);
#endif
        mbar_wait(bar_p + 16 * t + 8, j & 1, TAG_P_FULL);      // keys [64, 128)
        tc_fence_after();
        FA_TRACE_AT(2 + t, j, 3);
        if (elect_one_sync()) {
          issue_pv(t, vbuf, j > 0, FA_OPT_SPLITP ? kKStepsPV / 2 : 0, kKStepsPV);
          if (release) tc_commit(bar_empty + 8 * vbuf);
          if (last) {
            tc_commit(bar_o + 8 * t);
          } else {
            issue_s(t, qbuf, kbuf);
            tc_commit(bar_s + 8 * t);
            if (release) tc_commit(bar_empty + 8 * kbuf);
          }
        }
        __syncwarp();
        FA_TRACE_AT(2 + t, j, 6);
      };

      if (!split) {
        // ---- two Q tiles share every K/V tile: ring index of K_j is 2j, of V_j is 2j+1 ----
        if (n_tile[0] > 0) mbar_wait(bar_q, 0, TAG_Q_FULL);
        if (n_tile[1] > 0) mbar_wait(bar_q + 8, 0, TAG_Q_FULL);
        FA_TRACE_MISC(2, 1);
        wait_full(0);
        tc_fence_after();
        FA_TRACE_MISC(2, 2);
        if (elect_one_sync()) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (n_tile[t] > 0) {
              issue_s(t, t, 0);
              tc_commit(bar_s + 8 * t);
            }
          }
          tc_commit(bar_empty + 0);
        }
        __syncwarp();
        FA_TRACE_MISC(2, 3);
        for (int j = 0; j < n_max; ++j) {
          const int iv = 2 * j + 1, ik = 2 * j + 2;
          const int vbuf = iv % T::kNBuf, kbuf = ik % T::kNBuf;
          wait_full(iv);
          if (j + 1 < n_max) wait_full(ik);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < 2; ++t)
            if (j < n_tile[t]) step(t, j, j == n_tile[t] - 1, t, vbuf, kbuf, false);
          if (elect_one_sync()) {
            tc_commit(bar_empty + 8 * vbuf);
            if (j + 1 < n_max) tc_commit(bar_empty + 8 * kbuf);
          }
          __syncwarp();
        }
      } else {
        // ---- split-KV: each slot has its own K/V tiles; ring order K_A0, K_B0, then V_A(j), K_A(j+1), V_B(j), K_B(j+1) ----
        mbar_wait(bar_q, 0, TAG_Q_FULL);
        wait_full(0);
        if (n_tile[1] > 0) wait_full(1);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_s(0, 0, 0);
          tc_commit(bar_s);
          tc_commit(bar_empty);
          if (n_tile[1] > 0) {
            issue_s(1, 0, 1);
            tc_commit(bar_s + 8);
            tc_commit(bar_empty + 8);
          }
        }
        __syncwarp();
        int ring = n_tile[1] > 0 ? 2 : 1;
        for (int j = 0; j < n_max; ++j) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < n_tile[t]) {
              const bool last = (j == n_tile[t] - 1);
              const int iv = ring++;
              const int ik = last ? iv : ring++;
              wait_full(iv);
              if (!last) wait_full(ik);
              tc_fence_after();
              step(t, j, last, 0, iv % T::kNBuf, ik % T::kNBuf, true);
            }
          }
        }
      }
    }
  } else {
    // =========================== softmax + epilogue (warps 0-7) ===========================
    const int t = warp >> 2;                       // tile slot of this warpgroup
    const int r = (warp & 3) * 32 + lane;          // row within the tile == TMEM lane
    const int q_row = (t == 0 ? row0 : row_of_tile1) + r;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + T::kTmemS + t * kBlockN;
    const uint32_t tO = tmem_base + lane_base + T::kTmemO + t * kHeadDim;
    const int n_mine = n_tile[t];
    const float c = p.scale_log2;

    float m = -INFINITY;  // running (possibly stale) row max, in raw q.k units
    float l = 0.f;        // running row sum of exp2((s - m) * c)

    const bool tracer = (warp & 3) == 0 && lane == 0;
    (void)tracer;
    for (int j = 0; j < n_mine; ++j) {
      if (tracer) FA_TRACE_AT(t, j, 0);
      mbar_wait(bar_s + 8 * t, j & 1, TAG_S_FULL);
      tc_fence_after();
      if (tracer) FA_TRACE_AT(t, j, 1);
      float s[128];
      // masking: key kv0 + i is visible iff i <= limit
      const int kv0 = (kv_first[t] + j) * kBlockN;
      int limit = p.n_k - 1 - kv0;
      if (kCausal) limit = min(limit, q_row + p.causal_offset - kv0);
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#if FA_OPT_LDPIPE
      // chunk q's row-max pass runs while chunk q+1 is still in flight from TMEM
      tmem_ld32(tS, reinterpret_cast<uint32_t*>(&s[0]));
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        tc_wait_ld();
        if (q4 < 3) tmem_ld32(tS + (q4 + 1) * 32, reinterpret_cast<uint32_t*>(&s[(q4 + 1) * 32]));
        if (limit < kBlockN - 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q4 * 32 + i > limit) s[q4 * 32 + i] = -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, s[q4 * 32 + i]);
          mx1 = fmaxf(mx1, s[q4 * 32 + i + 1]);
          mx2 = fmaxf(mx2, s[q4 * 32 + i + 2]);
          mx3 = fmaxf(mx3, s[q4 * 32 + i + 3]);
        }
      }
#else
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld32(tS + q4 * 32, reinterpret_cast<uint32_t*>(&s[q4 * 32]));
      tc_wait_ld();
      if (limit < kBlockN - 1) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i > limit) s[i] = -INFINITY;
      }
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        mx0 = fmaxf(mx0, s[i]);
        mx1 = fmaxf(mx1, s[i + 1]);
        mx2 = fmaxf(mx2, s[i + 2]);
        mx3 = fmaxf(mx3, s[i + 3]);
      }
#endif
      const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
      if (tracer) FA_TRACE_AT(t, j, 2);

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
      // P = exp2(s*c - m*c) in two 64-key halves; each half is written to TMEM and handed to the MMA thread as soon
      // as it is complete, so the first half of P*V runs under the second half of the exps.
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = h * 64; i < h * 64 + 64; i += 4) {
          // which of these 4 elements take the polynomial route: the first FA_OPT_POLY/2 pairs of every 8 elements
          const bool kPoly01 = (FA_OPT_POLY >= 2) && ((i & 4) == 0);
          const bool kPoly23 = (FA_OPT_POLY >= 4) && ((i & 4) == 0);
#if FA_OPT_F2
          if constexpr (!kTF32) {
          float2 a01 = ffma2(make_float2(s[i], s[i + 1]), make_float2(c, c), make_float2(neg_mc, neg_mc));
          float2 a23 = ffma2(make_float2(s[i + 2], s[i + 3]), make_float2(c, c), make_float2(neg_mc, neg_mc));
          if (kPoly01) {
            a01 = exp2_poly2(a01);
            s[i] = a01.x;
            s[i + 1] = a01.y;
          } else {
            s[i] = ex2(a01.x);
            s[i + 1] = ex2(a01.y);
          }
          if (kPoly23) {
            a23 = exp2_poly2(a23);
            s[i + 2] = a23.x;
            s[i + 3] = a23.y;
          } else {
            s[i + 2] = ex2(a23.x);
            s[i + 3] = ex2(a23.y);
          }
          } else
#endif
          {
          if (kPoly01) {
            s[i] = exp2_poly(fmaf(s[i], c, neg_mc));
            s[i + 1] = exp2_poly(fmaf(s[i + 1], c, neg_mc));
          } else {
            s[i] = ex2(fmaf(s[i], c, neg_mc));
            s[i + 1] = ex2(fmaf(s[i + 1], c, neg_mc));
          }
          if (kPoly23) {
            s[i + 2] = exp2_poly(fmaf(s[i + 2], c, neg_mc));
            s[i + 3] = exp2_poly(fmaf(s[i + 3], c, neg_mc));
          } else {
            s[i + 2] = ex2(fmaf(s[i + 2], c, neg_mc));
            s[i + 3] = ex2(fmaf(s[i + 3], c, neg_mc));
          }
          }
          if constexpr (kTF32) {
            // kind::tf32 reads only the top 19 bits of P; sum exactly those values so that O = (sum P~ V) / (sum P~)
            // is normalised by what the tensor core actually multiplied (removes the truncation bias from O)
            s[i] = __uint_as_float(__float_as_uint(s[i]) & 0xFFFFE000u);
            s[i + 1] = __uint_as_float(__float_as_uint(s[i + 1]) & 0xFFFFE000u);
            s[i + 2] = __uint_as_float(__float_as_uint(s[i + 2]) & 0xFFFFE000u);
            s[i + 3] = __uint_as_float(__float_as_uint(s[i + 3]) & 0xFFFFE000u);
          }
#if FA_OPT_F2
          if constexpr (!kTF32) {
          const float2 s01 = fadd2(make_float2(l0, l1), make_float2(s[i], s[i + 1]));
          const float2 s23 = fadd2(make_float2(l2, l3), make_float2(s[i + 2], s[i + 3]));
          l0 = s01.x; l1 = s01.y; l2 = s23.x; l3 = s23.y;
          } else
#endif
          {
          l0 += s[i];
          l1 += s[i + 1];
          l2 += s[i + 2];
          l3 += s[i + 3];
          }
        }
        if constexpr (kTF32) {
          tmem_st32(tS + h * 64, reinterpret_cast<uint32_t*>(&s[h * 64]));
          tmem_st32(tS + h * 64 + 32, reinterpret_cast<uint32_t*>(&s[h * 64 + 32]));
        } else {
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) pk[i] = pack_bf16x2(s[h * 64 + 2 * i], s[h * 64 + 2 * i + 1]);
          tmem_st32(tS + h * 32, &pk[0]);
        }
#if FA_OPT_SPLITP
        if (tracer) FA_TRACE_AT(t, j, 3 + 2 * h);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p + 16 * t + 8 * h);
        if (tracer) FA_TRACE_AT(t, j, 4 + 2 * h);
#else
        if (h == 1) {
          if (tracer) FA_TRACE_AT(t, j, 5);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_p + 16 * t + 8);
          if (tracer) FA_TRACE_AT(t, j, 6);
        }
#endif
      }
      l += (l0 + l1) + (l2 + l3);
    }

    // ---- epilogue: O/l -> swizzled SMEM (reusing this tile's Q buffer) -> TMA store; LSE -> global ----
    if (tracer) FA_TRACE_MISC(t, 2);
    if (n_mine > 0) {
      mbar_wait(bar_o + 8 * t, 0, TAG_O_FINAL);
      tc_fence_after();
    }
    if (tracer) FA_TRACE_MISC(t, 3);
    // scale of this slot's accumulator and of the partner's partial (split-KV tail CTAs only) in the final O
    float f_self = (n_mine > 0 && l > 0.f) ? 1.0f / l : 0.f;
    float f_other = 0.f;
    float lse_val = -INFINITY;
    {
      const float m_safe = (m == -INFINITY) ? 0.f : m;
      if (n_mine > 0 && l > 0.f) lse_val = m_safe * p.scale + logf(l);
    }
    // split-KV exchange area: the K/V ring is dead once both slots' last MMAs have retired.
    // xO[col][row] fp32 (column-major: a warp's 32 rows are 32 consecutive words -> conflict-free), then m[128], l[128]
    const uint32_t xO = sKV;
    const uint32_t xML = sKV + kHeadDim * kBlockM * 4;
    bool stores = !(single_tile && t == 1);        // slot B of a 128-row CTA owns no output rows
    // both accumulators final => every tcgen05.mma that read the ring has retired
    if (split && n_tile[t ^ 1] > 0) mbar_wait(bar_o + 8 * (t ^ 1), 0, TAG_O_FINAL);
    if (split && n_tile[1] > 0) {
      tc_fence_after();
      if (t == 1) {
#pragma unroll
        for (int cc = 0; cc < kHeadDim / 32; ++cc) {
          uint32_t o[32];
          tmem_ld32(tO + cc * 32, o);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) st_shared_b32(xO + ((cc * 32 + i) * kBlockM + r) * 4, o[i]);
        }
        st_shared_b32(xML + r * 4, __float_as_uint(m));
        st_shared_b32(xML + (kBlockM + r) * 4, __float_as_uint(l));
        named_bar_arrive(kBarMerge, 256);
      } else {
        named_bar_sync(kBarMerge, 256);
        const float m_b = __uint_as_float(ld_shared_b32(xML + r * 4));
        const float l_b = __uint_as_float(ld_shared_b32(xML + (kBlockM + r) * 4));
        // log-sum-exp merge of (O_A, m, l) and (O_B, m_b, l_b); both maxima are in raw q.k units, sums in the exp2 domain
        const float m_all = fmaxf(m, m_b);
        const float a_a = (m == -INFINITY) ? 0.f : ex2((m - m_all) * c);
        const float a_b = (m_b == -INFINITY) ? 0.f : ex2((m_b - m_all) * c);
        const float l_all = l * a_a + l_b * a_b;
        const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
        f_self = a_a * inv;
        f_other = a_b * inv;
        lse_val = l_all > 0.f ? m_all * p.scale + logf(l_all) : -INFINITY;
      }
    }
    if (p.lse != nullptr && stores && q_row < p.n_q)
      p.lse[(static_cast<int64_t>(batch) * p.heads + head) * p.n_q + q_row] = lse_val;
    const uint32_t stage = sQ + t * T::kTileBytes;       // kDChunks boxes of 16 KB
    const uint32_t row_off = r * 128;
    const uint32_t sw = r & 7;
    constexpr int kRounds = T::kOChunks / T::kDChunks;   // 1, or 2 for bf16-in / fp32-out
    constexpr int kColsPerChunk = T::kOutElemsPerChunk;  // 32 (fp32) or 64 (bf16)
    if (stores) {
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
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f_self);
          if (split && n_tile[1] > 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              o[i] = __float_as_uint(fmaf(__uint_as_float(ld_shared_b32(xO + ((col0 + half * 32 + i) * kBlockM + r) * 4)), f_other,
                                          __uint_as_float(o[i])));
          }
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
      if (tracer) FA_TRACE_MISC(t, 4);
      if ((warp & 3) == 0 && lane == 0) {
#pragma unroll
        for (int ch = 0; ch < T::kDChunks; ++ch)
          tma_store_4d(&tm_o, stage + ch * kChunkBytes, (round * T::kDChunks + ch) * kColsPerChunk, row0 + t * kBlockM, head,
                       batch);
        tma_store_commit();
        tma_store_wait_read();
        FA_TRACE_MISC(t, 5);
      }
      if (round + 1 < kRounds) named_bar_sync(1 + t, 128);
    }
    if ((warp & 3) == 0 && lane == 0) {
      tma_store_wait_all();
      FA_TRACE_MISC(t, 6);
    }
    }
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
