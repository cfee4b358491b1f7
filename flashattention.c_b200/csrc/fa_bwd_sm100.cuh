// fa_bwd_sm100.cuh — FlashAttention backward for sm_100a (16-bit operands, head dim <= 128): dQ, dK, dV of
// O = softmax(scale Q K^T [+ causal]) V from (Q, K, V, O, LSE, dO).  The reference is forward only (README.md:33 lists what
// it leaves open); SURVEY §8 (f4) names the backward pass as the widening step after the forward rows.
//
// Two launches of ONE kernel template, both atomics-free and deterministic:
//   kDKV = true   a CTA owns 128 keys of one K/V head (resident tiles K_j, V_j) and streams 128-row tiles of (Q, dO) of every
//                 query head of its group:   S^T = K Q^T,  dP^T = V dO^T   (tcgen05.mma, A and B from SMEM, D in TMEM)
//                 P^T = exp2(S^T c - LSE2[q]),  dS^T = P^T (dP^T - D[q])   (two threads per key row, statistics per column)
//                 dV += P^T dO,  dK += dS^T Q                              (A = P^T / dS^T read straight from TMEM)
//   kDKV = false  a CTA owns 128 query rows of one head (resident Q_i, dO_i) and streams 128-key tiles of (K, V):
//                 S = Q K^T,  dP = dO V^T,  dS = P (dP - D[q])  (statistics per row, in registers),  dQ += dS K
// i.e. 4 + 3 = 7 contractions instead of the 5 of a single-pass backward: recomputing S and dP in the second launch is what
// buys the absence of a dQ reduction across CTAs (no atomics, no dQ workspace, bit-reproducible results), and every
// contraction has exactly the operand forms the forward kernel uses (K-major SMEM x K-major SMEM; TMEM x MN-major SMEM), on
// the tiles as TMA delivers them (SWIZZLE_128B boxes of 128 bytes x rows): a streamed tile is read K-major by the first two
// contractions and MN-major by the accumulating ones.
//
// D[q] = rowsum(dO * O) and LSE2[q] = LSE * log2(e) come from fa_bwd_prep_kernel in a workspace whose rows are padded to a
// multiple of 128 per (batch, head) (+inf / 0 in the padding and for rows that saw no key: P = exp2(x - inf) = 0 there).
//
// Pipeline: warp 8 lane 0 = TMA producer (one ring per streamed tensor), warp 9 lane 0 = MMA issuer, warps 0-7 = two threads per
// TMEM lane (warp w: lanes 32 (w % 4) .. + 31, columns 64 (w / 4) .. + 63 of S and dP).  Streamed tiles are 128 rows: every
// contraction is then a run of eight 128 x 128 x 16 MMAs at the full tensor rate (64-row tiles were measured first: their
// 128 x 64 MMAs cost 54 instead of 33.5 cycles per 64 columns, and the fixed costs per step — barrier waits, the switch between
// SMEM- and TMEM-sourced MMAs — weigh twice as much: profiles/r02_bwd_trace_v2_8warps.txt).  TMEM is exactly full at d = 128
// (S 128 + dP 128 + two accumulators), so S and dP cannot be double-buffered; instead the two regions alternate: the tensor
// pipe runs   [P(i)] dV(i), S(i+1), [dS(i)] dK(i), dP(i+1)   — while the warps compute P(i) from S(i) it runs dK(i-1) and dP(i),
// while they compute dS(i) from dP(i) it runs dV(i) and S(i+1).  P^T / dS^T (packed 16-bit pairs) overwrite the first half of
// each thread's own S / dP columns.
// The scale of dS (dS_raw = scale * P (dP - D)) is applied once, to the finished dQ / dK accumulators.
#pragma once
#include <type_traits>

#include "fa_simt.cuh"   // ld_as_float
#include "ptx.cuh"

// -DFA_BWD_TRACE=1: CTA (0, 0, 0) records clock64() at the pipeline hand-offs of its first 32 steps into BwdParams::trace
// ([role 0..3][step][slot 0..7]; roles: compute warp 0, compute warp 4, MMA thread, producer)
#ifndef FA_BWD_TRACE
#define FA_BWD_TRACE 0
#endif
#if FA_BWD_TRACE
#define FA_BWD_TRACE_AT(role, step, slot)                                                                       \
  do {                                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (step) < 32)             \
      p.trace[((role) * 32 + (step)) * 8 + (slot)] = static_cast<unsigned long long>(clock64());                \
  } while (0)
#else
#define FA_BWD_TRACE_AT(role, step, slot) do { } while (0)
#endif

// Of every FA_BWD_POLY_DEN consecutive element pairs of P, the first FA_BWD_POLY_NUM get exp2 from the FMA-pipe polynomial
// (ptx.cuh: exp2_poly2, relative error 7.5e-5 — P is rounded to 16 bits right after) instead of MUFU.EX2: both launches evaluate
// every P element, 16384 per 128 x 128 step at 16 per clock, and the traces show the exp pass, not the tensor pipe, setting the
// step time (profiles/r02_bwd_trace_v3_stream128.txt).
// Per launch, by ncu durations (profiles/r02_bwd_ab_poly_fraction_ncu.log): the dK/dV launch is fastest at 1/2, the dQ launch
// (which has no P to pack and store, so fewer instructions per element to begin with) at 1/4 - 3/8.
#ifndef FA_BWD_POLY_NUM
#define FA_BWD_POLY_NUM 1
#endif
#ifndef FA_BWD_POLY_DEN
#define FA_BWD_POLY_DEN 2
#endif
#ifndef FA_BWD_POLY_NUM_DQ
#define FA_BWD_POLY_NUM_DQ 3
#endif
#ifndef FA_BWD_POLY_DEN_DQ
#define FA_BWD_POLY_DEN_DQ 8
#endif

namespace fa {

struct BwdParams {
  float scale;        // multiplies q.k (as in the forward)
  float scale_log2;   // scale * log2(e)
  int n_q, n_k, heads, kv_heads, batch;
  int kv_group;       // query heads per K/V head
  int head_dim;       // true head dim (<= the instance's; columns beyond it are zero in SMEM and never stored)
  int causal, causal_offset;   // key j visible to row i iff j <= i + causal_offset (n_k - n_q)
  int n_q_pad;        // row pitch of the statistics workspace per (batch, head): n_q rounded up to 128
  const float* l2;    // [batch, heads, n_q_pad]
  const float* dsum;  // [batch, heads, n_q_pad]
  void* out0;         // kDKV: dV [batch, kv_heads, n_k, d]   else: dQ [batch, heads, n_q, d]   (element strides below)
  int64_t o0_sb, o0_sh, o0_sn;
  void* out1;         // kDKV: dK
  int64_t o1_sb, o1_sh, o1_sn;
  unsigned long long* trace;   // FA_BWD_TRACE builds only; nullptr otherwise
};

// threads per TMEM lane in the element-wise pass (2: 8 warps, 64 columns of S / dP per thread; 4: 16 warps, 32 columns)
#ifndef FA_BWD_SPLIT
#define FA_BWD_SPLIT 2
#endif
constexpr int kBwdSplit = FA_BWD_SPLIT;
static_assert(kBwdSplit == 2 || kBwdSplit == 4, "two or four threads per TMEM lane");
constexpr int kBwdThreads = (4 * kBwdSplit + 2) * 32;   // compute warps + producer warp + MMA warp
constexpr int kBwdProdWarp = 4 * kBwdSplit, kBwdMmaWarp = 4 * kBwdSplit + 1;
constexpr bool kBwdStatRegs = kBwdSplit == 2;            // per-column statistics held in registers (else read from SMEM where used)
constexpr int kBwdRes = 128;       // rows of a resident tile (keys of the dK/dV launch, query rows of the dQ launch)
constexpr int kBwdStr = 128;       // rows of a streamed tile
constexpr int kBwdHalf = 128 / kBwdSplit;   // S / dP columns per compute thread

template <int kHeadDim>
struct BwdTraits {
  static_assert(kHeadDim == 64 || kHeadDim == 128, "backward instances: 128- and 256-byte rows of 16-bit elements");
  static constexpr int kDChunks = kHeadDim * 2 / 128;            // 128-byte column chunks per row
  static constexpr int kChunkBytes = 128 * 128;                  // one TMA box: 128 rows x 128 bytes
  static constexpr int kTileBytes = kDChunks * kChunkBytes;      // resident and streamed tiles alike
  // The two streamed tensors have rings of their own: a tile of the first (Q in the dK/dV launch, K in the dQ launch) is read by
  // the step's first contraction (S) and by its last (dK / dQ), and its successor-but-one is needed only one contraction after
  // that — with two buffers its TMA load sat on the critical path (profiles/r02_bwd_trace_v4_poly.txt) — so it gets three; a tile
  // of the second (dO / V) is done with much earlier and two are enough.  At d = 128 that is 64 + 96 + 64 KB: all of SMEM.
  static constexpr int kRing1 = kDChunks == 1 ? 4 : 3;
  static constexpr int kRing2 = kDChunks == 1 ? 3 : 2;
  // D[i] rides in dO_i's ring slot and is read by the element-wise warps.  With a two-deep ring (d = 128) the slot must go back
  // to the producer right behind dV(i), so the warps take D[i] into registers BEFORE they hand P(i) over (which is what lets
  // dV(i) be issued); with three slots (d = 64) they read it after the hand-over — off the P chain — and the slot is released
  // one contraction later, behind dK(i) (ncu durations, profiles/r02_bwd_ab_stat_release_ncu.log: each choice is 4-5 % faster
  // than the other on its instance).
  static constexpr bool kEarlyD = kBwdStatRegs && kRing2 == 2;
  static constexpr int kStatBytes = kBwdStr * 4;                 // LSE2[128] (rides with ring 1) or D[128] (ring 2) of a (Q, dO) tile
  static constexpr int kNumBarriers = 1 + 2 * kRing1 + 2 * kRing2 + 4 + 1;
  // (the dynamic SMEM window is 1024-byte aligned — checked at kernel entry — so there is no alignment slack: there is no room for it)
  static constexpr int kSmemBytes = (2 + kRing1 + kRing2) * kTileBytes + (kRing1 + kRing2) * kStatBytes + kNumBarriers * 8 + 16;
  static constexpr int kTmemS = 0, kTmemDP = 128;                // S and dP: 128 fp32 columns each (P / dS alias them)
  static constexpr int kTmemAcc = 256;                           // acc0 at 256, acc1 at 256 + kHeadDim
  static_assert(kTmemAcc + 2 * kHeadDim <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 227 * 1024, "SMEM budget");
};

enum : uint32_t { TAG_B_RES = 21, TAG_B_FULL = 22, TAG_B_EMPTY = 23, TAG_B_S = 24, TAG_B_P = 25, TAG_B_ACC = 26, TAG_B_DP = 27, TAG_B_DS = 28 };

// plain (non-tensor) bulk copy global -> shared, completing on an mbarrier; 16-byte aligned on both sides, bytes % 16 == 0
FA_DEVINL void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}

template <int kHeadDim, bool kF16, bool kDKV>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_r1, const __grid_constant__ CUtensorMap tm_r2,
                    const __grid_constant__ CUtensorMap tm_t1, const __grid_constant__ CUtensorMap tm_t2, const BwdParams p) {
  using T = BwdTraits<kHeadDim>;
  extern __shared__ __align__(1024) uint8_t bwd_smem_raw[];
  const uint32_t smem0 = smem_u32(bwd_smem_raw);
  if ((smem0 & 1023u) != 0u) __trap();                             // the SWIZZLE_128B tiles need it (see BwdTraits::kSmemBytes)
  const uint32_t sR1 = smem0;                                      // kDKV: K_j    else: Q_i
  const uint32_t sR2 = sR1 + T::kTileBytes;                        // kDKV: V_j    else: dO_i
  const uint32_t sT1 = sR2 + T::kTileBytes;                        // [kRing1]  kDKV: Q    else: K
  const uint32_t sT2 = sT1 + T::kRing1 * T::kTileBytes;            // [kRing2]  kDKV: dO   else: V
  const uint32_t sL2 = sT2 + T::kRing2 * T::kTileBytes;            // [kRing1][128]  LSE2 of the Q tile (kDKV)
  const uint32_t sD = sL2 + T::kRing1 * T::kStatBytes;             // [kRing2][128]  D of the dO tile (kDKV)
  const uint32_t bar_res = sD + T::kRing2 * T::kStatBytes;
  const uint32_t bar_full1 = bar_res + 8;                          // [kRing1]
  const uint32_t bar_empty1 = bar_full1 + 8 * T::kRing1;           // [kRing1]
  const uint32_t bar_full2 = bar_empty1 + 8 * T::kRing1;           // [kRing2]
  const uint32_t bar_empty2 = bar_full2 + 8 * T::kRing2;           // [kRing2]
  const uint32_t bar_s = bar_empty2 + 8 * T::kRing2;               // S of the step complete
  const uint32_t bar_dp = bar_s + 8;                               // dP of the step complete
  const uint32_t bar_p = bar_dp + 8;                               // kDKV: P written over S   else: S is in registers (its columns are free)
  const uint32_t bar_ds = bar_p + 8;                               // dS written over dP
  const uint32_t bar_acc = bar_ds + 8;                             // accumulators final
  const uint32_t s_tmem_ptr = bar_acc + 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int batch = blockIdx.z;
  const int head_r = blockIdx.y;                                   // head of the resident tiles (kDKV: a K/V head)
  // the dQ launch takes its causal tiles heaviest (= last rows) first
  const int tile = (!kDKV && p.causal) ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int row0 = tile * kBwdRes;

  // ---- the streamed steps of this CTA ----
  // kDKV: step -> (g, i): query head head_r * kv_group + g, rows [128 i, 128 i + 128), i from the first tile with a row that
  //       sees one of this CTA's keys.  else: step j -> keys [128 j, 128 j + 128), up to the last key any of the CTA's rows sees.
  int i_first = 0, per_head = 0, n_steps = 0;
  if (kDKV) {
    const int nq_t = (p.n_q + kBwdStr - 1) / kBwdStr;
    if (p.causal) i_first = max(0, row0 - p.causal_offset) / kBwdStr;
    per_head = max(0, nq_t - i_first);
    n_steps = per_head * p.kv_group;
  } else {
    int last = p.n_k - 1;
    if (p.causal) last = min(last, min(row0 + kBwdRes - 1, p.n_q - 1) + p.causal_offset);
    n_steps = last < 0 ? 0 : last / kBwdStr + 1;
  }

  if (warp == kBwdMmaWarp && lane == 0) {
    mbar_init(bar_res, 1);
    for (int i = 0; i < T::kRing1; ++i) {
      mbar_init(bar_full1 + 8 * i, 1);
      mbar_init(bar_empty1 + 8 * i, 1);
    }
    for (int i = 0; i < T::kRing2; ++i) {
      mbar_init(bar_full2 + 8 * i, 1);
      mbar_init(bar_empty2 + 8 * i, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_p, 128 * kBwdSplit);
    mbar_init(bar_ds, 128 * kBwdSplit);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == kBwdProdWarp) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_r1);
      tma_prefetch_desc(&tm_r2);
      tma_prefetch_desc(&tm_t1);
      tma_prefetch_desc(&tm_t2);
    }
    tmem_alloc(s_tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem_ptr));

  if (warp == kBwdProdWarp) {
    // =========================== TMA producer ===========================
    if (lane == 0 && n_steps > 0) {
      mbar_arrive_expect_tx(bar_res, 2 * T::kTileBytes);
#pragma unroll
      for (int c = 0; c < T::kDChunks; ++c) {
        tma_load_4d(sR1 + c * T::kChunkBytes, &tm_r1, bar_res, c * 64, row0, head_r, batch);
        tma_load_4d(sR2 + c * T::kChunkBytes, &tm_r2, bar_res, c * 64, row0, head_r, batch);
      }
      for (int step = 0; step < n_steps; ++step) {
        int head_t, srow;
        if (kDKV) {
          head_t = head_r * p.kv_group + step / per_head;
          srow = (i_first + step % per_head) * kBwdStr;
        } else {
          head_t = head_r / p.kv_group;
          srow = step * kBwdStr;
        }
        const int64_t stat_off = (static_cast<int64_t>(batch) * p.heads + head_t) * p.n_q_pad + srow;
        {
          const int s1 = step % T::kRing1;
          if (step >= T::kRing1) mbar_wait(bar_empty1 + 8 * s1, ((step / T::kRing1) - 1) & 1, TAG_B_EMPTY);
          FA_BWD_TRACE_AT(3, step, 0);
          mbar_arrive_expect_tx(bar_full1 + 8 * s1, T::kTileBytes + (kDKV ? T::kStatBytes : 0));
#pragma unroll
          for (int c = 0; c < T::kDChunks; ++c)
            tma_load_4d(sT1 + s1 * T::kTileBytes + c * T::kChunkBytes, &tm_t1, bar_full1 + 8 * s1, c * 64, srow, head_t, batch);
          if (kDKV) bulk_load(sL2 + s1 * T::kStatBytes, p.l2 + stat_off, T::kStatBytes, bar_full1 + 8 * s1);
        }
        {
          const int s2 = step % T::kRing2;
          if (step >= T::kRing2) mbar_wait(bar_empty2 + 8 * s2, ((step / T::kRing2) - 1) & 1, TAG_B_EMPTY);
          FA_BWD_TRACE_AT(3, step, 1);
          mbar_arrive_expect_tx(bar_full2 + 8 * s2, T::kTileBytes + (kDKV ? T::kStatBytes : 0));
#pragma unroll
          for (int c = 0; c < T::kDChunks; ++c)
            tma_load_4d(sT2 + s2 * T::kTileBytes + c * T::kChunkBytes, &tm_t2, bar_full2 + 8 * s2, c * 64, srow, head_t, batch);
          if (kDKV) bulk_load(sD + s2 * T::kStatBytes, p.dsum + stat_off, T::kStatBytes, bar_full2 + 8 * s2);
        }
      }
    }
  } else if (warp == kBwdMmaWarp) {
    // =========================== MMA issuer ===========================
    // Order on the (in-order) tensor pipe:  S(0) dP(0) | then per step i:  [P(i)] dV(i)  S(i+1)  [dS(i)] dK(i)  dP(i+1)
    // (dQ launch: no dV; dK -> dQ).  S(i+1) overwrites the S columns once dV(i) has read P(i) out of them, dP(i+1) the dP columns
    // once dK(i) has read dS(i): the two 128-column regions are the double buffer, one contraction apart.
    // The whole warp follows the (warp-uniform) control flow and waits on the mbarriers; one elected lane issues.  (Issued from
    // inside an `if (lane == 0)` region every tcgen05.mma is wrapped in an ELECT / branch loop of its own by the compiler — the
    // operands of a warp-uniform instruction could differ between lanes for all it knows — and a step's 32 MMAs then issue more
    // slowly than the tensor pipe retires them: profiles/r02_mma_probe_bwd_pattern.log has the pipe alone at 2078 cycles per step
    // against 2700-2900 measured in the kernel.)
    if (n_steps > 0) {
      constexpr uint32_t kFmt = kF16 ? 0u : 1u;
      constexpr uint32_t idesc_sd = make_idesc(kFmt, 0, kBwdRes, kBwdStr);      // S, dP: 128 x 128, both operands K-major
      constexpr uint32_t idesc_acc = make_idesc(kFmt, 1, kBwdRes, kHeadDim);    // accumulators: 128 x d, B MN-major
      constexpr uint64_t hi_kmajor = make_sdesc_hi_sw128(16, 1024);
      // a streamed tile as the MN-major B operand (N = d, K = its 128 rows): LBO = stride between the 128-byte column chunks,
      // SBO = stride between 8-row groups
      constexpr uint64_t hi_mnmajor = make_sdesc_hi_sw128(T::kChunkBytes, 1024);
      constexpr int kKStepsD = kHeadDim / 16;       // k-steps over the head dim (S, dP)
      constexpr int kKStepsR = kBwdStr / 16;        // k-steps over the streamed rows (accumulators)
      mbar_wait(bar_res, 0, TAG_B_RES);
      tc_fence_after();
      const uint64_t r1d = sdesc_at(hi_kmajor, sR1);
      const uint64_t r2d = sdesc_at(hi_kmajor, sR2);
      const uint32_t tS = tmem_base + T::kTmemS, tDP = tmem_base + T::kTmemDP;
      const uint32_t acc0 = tmem_base + T::kTmemAcc, acc1 = acc0 + kHeadDim;
      // SMEM address of the step's tile of streamed tensor `which` (0: ring 1, 1: ring 2)
      auto tile_of = [&](int step, int which) {
        return which == 0 ? sT1 + (step % T::kRing1) * T::kTileBytes : sT2 + (step % T::kRing2) * T::kTileBytes;
      };
      auto wait_full = [&](int step, int which) {
        if (which == 0) mbar_wait(bar_full1 + 8 * (step % T::kRing1), (step / T::kRing1) & 1, TAG_B_FULL);
        else mbar_wait(bar_full2 + 8 * (step % T::kRing2), (step / T::kRing2) & 1, TAG_B_FULL);
        tc_fence_after();
      };
      // D[tmem d] = R (resident, K-major) x T^T (the step's tile of streamed tensor `which`, K-major)
      auto issue_rt = [&](uint32_t d, uint64_t rd, int step, int which) {
        const uint64_t td = sdesc_at(hi_kmajor, tile_of(step, which));
#pragma unroll
        for (int kk = 0; kk < kKStepsD; ++kk) {
          const uint32_t off = ((kk >> 2) * T::kChunkBytes + (kk & 3) * 32) >> 4;
          mma_ss<false>(d, rd + off, td + off, idesc_sd, kk > 0 ? 1u : 0u);
        }
      };
      // acc += A (TMEM: packed P or dS; the streamed rows [kBwdHalf h, kBwdHalf (h + 1)) of thread group h sit in the first half of
      // that group's own columns of the region) x T (MN-major)
      auto issue_acc = [&](uint32_t acc, uint32_t a, int step, int which) {
        const uint64_t tm = sdesc_at(hi_mnmajor, tile_of(step, which));
#pragma unroll
        for (int ks = 0; ks < kKStepsR; ++ks)
          mma_ts<false>(acc, a + static_cast<uint32_t>((ks / (kBwdHalf / 16)) * kBwdHalf + (ks % (kBwdHalf / 16)) * 8),
                        tm + static_cast<uint32_t>(ks * 128), idesc_acc,
                        (step > 0 || ks > 0) ? 1u : 0u);
      };
      wait_full(0, 0);
      if (elect_one_sync()) {
        issue_rt(tS, r1d, 0, 0);
        tc_commit(bar_s);
      }
      __syncwarp();
      wait_full(0, 1);
      if (elect_one_sync()) {
        issue_rt(tDP, r2d, 0, 1);
        tc_commit(bar_dp);
        if (!kDKV) tc_commit(bar_empty2);                  // dQ launch: V_0 is only read by dP(0)
      }
      __syncwarp();
      for (int step = 0; step < n_steps; ++step) {
        const uint32_t par = static_cast<uint32_t>(step & 1);
        const bool more = step + 1 < n_steps;
        if (lane == 0) FA_BWD_TRACE_AT(2, step, 0);
        // the next step's tiles were requested one to two steps ago: check for them now, while the pipe is still busy with what
        // was issued last, so that nothing stands between a P / dS hand-over and the issue that waits for it
        if (more) {
          wait_full(step + 1, 0);
          wait_full(step + 1, 1);
        }
        mbar_wait(bar_p, par, TAG_B_P);
        tc_fence_after();
        if (lane == 0) FA_BWD_TRACE_AT(2, step, 1);
        if (elect_one_sync()) {
          if (kDKV) {
            issue_acc(acc0, tS, step, 1);                    // dV += P^T dO
            // dO_i has been read by dP(i) and dV(i).  D[i] rides in the same ring slot and is read by the element-wise warps: with
            // kEarlyD they have it in registers before they hand P(i) over, and the slot goes back to the producer behind dV(i)
            if (T::kEarlyD) tc_commit(bar_empty2 + 8 * (step % T::kRing2));
          }
          if (more) {
            issue_rt(tS, r1d, step + 1, 0);                  // S of the next step
            tc_commit(bar_s);
          }
        }
        __syncwarp();
        if (lane == 0) FA_BWD_TRACE_AT(2, step, 2);
        mbar_wait(bar_ds, par, TAG_B_DS);
        tc_fence_after();
        if (lane == 0) FA_BWD_TRACE_AT(2, step, 3);
        if (elect_one_sync()) {
          issue_acc(kDKV ? acc1 : acc0, tDP, step, 0);       // dK += dS^T Q   /   dQ += dS K
          tc_commit(bar_empty1 + 8 * (step % T::kRing1));    // Q_i / K_j has been read by S and by dK / dQ once everything issued so far completes
          // (!kEarlyD: D[i] is read after the P hand-over, possibly after dV(i) has been issued, so dO_i's slot is released behind
          // dK(i), which needs dS(i), which is written after D[i] has been read)
          if (kDKV && !T::kEarlyD) tc_commit(bar_empty2 + 8 * (step % T::kRing2));
          if (more) {
            issue_rt(tDP, r2d, step + 1, 1);                 // dP of the next step
            tc_commit(bar_dp);
            if (!kDKV) tc_commit(bar_empty2 + 8 * ((step + 1) % T::kRing2));   // dQ launch: V_j is only read by dP(j)
          }
        }
        __syncwarp();
        if (lane == 0) FA_BWD_TRACE_AT(2, step, 4);
      }
      if (elect_one_sync()) tc_commit(bar_acc);
      __syncwarp();
    }
  } else {
    // =========================== P, dS (two threads per TMEM lane) + epilogue ===========================
    const int r = (warp & 3) * 32 + lane;
    const int hh = warp >> 2;             // which half of the 128 columns of a step (and of the accumulator columns in the epilogue)
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int my_row = row0 + r;          // kDKV: key index    else: query row
    float l2r = INFINITY, dr = 0.f;       // dQ launch: the row's statistics
    if (!kDKV) {
      const int64_t off = (static_cast<int64_t>(batch) * p.heads + head_r) * p.n_q_pad + my_row;   // my_row < n_q_pad always
      l2r = p.l2[off];
      dr = p.dsum[off];
    }
    const bool row_valid = kDKV ? (my_row < p.n_k) : true;   // (query rows beyond n_q have LSE2 = +inf)
    const uint32_t tS = tmem_base + lane_base + T::kTmemS + hh * kBwdHalf;     // this thread's 64 columns of S ...
    const uint32_t tDP = tmem_base + lane_base + T::kTmemDP + hh * kBwdHalf;   // ... and of dP
    const bool tracer = FA_BWD_TRACE && (warp & 3) == 0 && lane == 0 && hh < 2;
    for (int step = 0; step < n_steps; ++step) {
      const int s1 = step % T::kRing1, s2 = step % T::kRing2;
      const uint32_t par = static_cast<uint32_t>(step & 1);
      // column c (0 .. 127) of this step is visible to this thread's row iff c_lo <= c <= c_hi; then in its own numbering (0 .. 63)
      int c_lo = 0, c_hi = kBwdStr - 1;
      if (kDKV) {
        const int q0 = (i_first + step % per_head) * kBwdStr;
        c_hi = p.n_q - 1 - q0;                                         // (rows beyond n_q: LSE2 = +inf, but the polynomial exp2 clamps)
        if (p.causal) c_lo = my_row - p.causal_offset - q0;            // q0 + c + offset >= key
        if (!row_valid) c_lo = kBwdStr;
      } else {
        const int k0 = step * kBwdStr;
        c_hi = p.n_k - 1 - k0;
        if (p.causal) c_hi = min(c_hi, my_row + p.causal_offset - k0);
        if (l2r == INFINITY) c_hi = -1;                                // a row beyond n_q, or one that sees no key at all
      }
      c_lo -= hh * kBwdHalf;
      c_hi -= hh * kBwdHalf;
      const bool masked = c_lo > 0 || c_hi < kBwdHalf - 1;
      const uint32_t s_l2 = sL2 + s1 * T::kStatBytes + hh * kBwdHalf * 4;
      const uint32_t s_d = sD + s2 * T::kStatBytes + hh * kBwdHalf * 4;
      // the statistics of this thread's columns: in registers, loaded before S is waited for (LSE2 now, D after the exps), or —
      // with four threads per lane, where the register file has no room for them — read from SMEM where they are used
      float stat[kBwdStatRegs ? kBwdHalf : 4];
      auto load_stats = [&](uint32_t addr) {
        if constexpr (kBwdStatRegs) {
#pragma unroll
          for (int c4 = 0; c4 < kBwdHalf / 4; ++c4) {
            uint32_t a0, a1, a2, a3;
            ld_shared_v4(addr + c4 * 16, a0, a1, a2, a3);
            stat[c4 * 4] = __uint_as_float(a0); stat[c4 * 4 + 1] = __uint_as_float(a1);
            stat[c4 * 4 + 2] = __uint_as_float(a2); stat[c4 * 4 + 3] = __uint_as_float(a3);
          }
        }
      };
      // statistics of columns 4 c4 .. 4 c4 + 3
      auto stat_quad = [&](uint32_t addr, int c4, float (&out)[4]) {
        if constexpr (kBwdStatRegs) {
#pragma unroll
          for (int u = 0; u < 4; ++u) out[u] = stat[c4 * 4 + u];
        } else {
          uint32_t a0, a1, a2, a3;
          ld_shared_v4(addr + c4 * 16, a0, a1, a2, a3);
          out[0] = __uint_as_float(a0); out[1] = __uint_as_float(a1); out[2] = __uint_as_float(a2); out[3] = __uint_as_float(a3);
        }
      };
      if (kDKV) {
        mbar_wait(bar_full1 + 8 * s1, (step / T::kRing1) & 1, TAG_B_FULL);
        load_stats(s_l2);
      }

      // ---- P = exp2(S c - LSE2) ----
      mbar_wait(bar_s, par, TAG_B_S);
      tc_fence_after();
      if (tracer) FA_BWD_TRACE_AT(hh, step, 0);
      float pr[kBwdHalf];
#pragma unroll
      for (int q = 0; q < kBwdHalf / 32; ++q) tmem_ld32(tS + q * 32, reinterpret_cast<uint32_t*>(&pr[q * 32]));
      tc_wait_ld();
      if (!kDKV) {   // S is in registers and nothing is written back over it: the next S may be issued
        tc_fence_before();
        mbar_arrive(bar_p);
      }
      // two copies of the pass: the masked one (a diagonal or tail tile: two compares and two selects per element) only when some
      // thread of the warp needs it
      auto exp_pass = [&](auto with_mask) {
#pragma unroll
        for (int c4 = 0; c4 < kBwdHalf / 4; ++c4) {
          float lq[4] = {l2r, l2r, l2r, l2r};
          if (kDKV) stat_quad(s_l2, c4, lq);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = c4 * 2 + u;   // pair index: columns 2 i, 2 i + 1
            float2 e = ffma2(make_float2(pr[2 * i], pr[2 * i + 1]), make_float2(p.scale_log2, p.scale_log2), make_float2(-lq[2 * u], -lq[2 * u + 1]));
            constexpr int kPolyNum = kDKV ? FA_BWD_POLY_NUM : FA_BWD_POLY_NUM_DQ, kPolyDen = kDKV ? FA_BWD_POLY_DEN : FA_BWD_POLY_DEN_DQ;
            if ((i % kPolyDen) < kPolyNum) {
              e = exp2_poly2(e);
            } else {
              e.x = ex2(e.x);
              e.y = ex2(e.y);
            }
            if constexpr (decltype(with_mask)::value) {
              if (2 * i < c_lo || 2 * i > c_hi) e.x = 0.f;
              if (2 * i + 1 < c_lo || 2 * i + 1 > c_hi) e.y = 0.f;
            }
            pr[2 * i] = e.x;
            pr[2 * i + 1] = e.y;
          }
        }
      };
      if (__any_sync(0xffffffffu, masked)) exp_pass(std::true_type{});
      else exp_pass(std::false_type{});
      if (tracer) FA_BWD_TRACE_AT(hh, step, 1);
      if (kDKV) {
        // kEarlyD: D[i] into the registers LSE2[i] has just left — BEFORE P(i) is handed over: the hand-over lets the MMA warp issue dV(i),
        // the last reader of dO_i on the tensor pipe, behind which the ring slot that also holds D[i] returns to the producer.
        // (Read after the hand-over, a slow warp could find the statistics of step i + 2 in the slot: correct at full speed,
        // wrong dK under compute-sanitizer, which stretches such windows — tests/test_gpu_backward.py runs that check.)
        // The loads complete under the packing and the TMEM store of P; the empty asm below makes the arrive wait for them.
        if (T::kEarlyD) {
          mbar_wait(bar_full2 + 8 * s2, (step / T::kRing2) & 1, TAG_B_FULL);
          load_stats(s_d);
        }
        // P^T (two 16-bit values per column) goes over the first half of the thread's OWN columns of S, which it has in registers
        uint32_t pk[kBwdHalf / 2];
#pragma unroll
        for (int i = 0; i < kBwdHalf / 2; ++i) pk[i] = pack_16x2<kF16>(pr[2 * i], pr[2 * i + 1]);
        if constexpr (kBwdHalf == 64) tmem_st32(tS, pk);
        else tmem_st16(tS, pk);
        tc_wait_st();
        tc_fence_before();
        if constexpr (T::kEarlyD) {
#pragma unroll
          for (int i = 0; i < kBwdHalf; ++i) asm volatile("" ::"f"(stat[i]) : "memory");   // (keeps the loads ahead of the arrive)
        }
        mbar_arrive(bar_p);
      }
      if (tracer) FA_BWD_TRACE_AT(hh, step, 2);
      if (kDKV && !T::kEarlyD) {
        mbar_wait(bar_full2 + 8 * s2, (step / T::kRing2) & 1, TAG_B_FULL);
        load_stats(s_d);
      }

      // ---- dS = P (dP - D) ----
      mbar_wait(bar_dp, par, TAG_B_DP);
      tc_fence_after();
      if (tracer) FA_BWD_TRACE_AT(hh, step, 3);
      uint32_t dk[kBwdHalf / 2];
#pragma unroll
      for (int h2 = 0; h2 < kBwdHalf / 32; ++h2) {   // in pieces of 32 columns: 32 live dP registers
        float dp[32];
        tmem_ld32(tDP + h2 * 32, reinterpret_cast<uint32_t*>(&dp[0]));
        tc_wait_ld();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float dq[4] = {dr, dr, dr, dr};
          if (kDKV) stat_quad(s_d, h2 * 8 + c4, dq);
#pragma unroll
          for (int u = 0; u < 2; ++u) {   // packed FADD2 / FMUL2: one instruction per two elements
            const int c = h2 * 32 + c4 * 4 + 2 * u;
            const float2 dv = fmul2(make_float2(pr[c], pr[c + 1]),
                                    fadd2(make_float2(dp[c4 * 4 + 2 * u], dp[c4 * 4 + 2 * u + 1]), make_float2(-dq[2 * u], -dq[2 * u + 1])));
            dk[h2 * 16 + c4 * 2 + u] = pack_16x2<kF16>(dv.x, dv.y);
          }
        }
      }
      // dS over the first half of the thread's own columns of dP (all of them have been read by now)
      if constexpr (kBwdHalf == 64) tmem_st32(tDP, dk);
      else tmem_st16(tDP, dk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_ds);
      if (tracer) FA_BWD_TRACE_AT(hh, step, 4);
    }

    // ---- epilogue: accumulators -> (scale) -> 16-bit -> global, one row per pair of threads ----
    const bool have = n_steps > 0;
    if (have) {
      mbar_wait(bar_acc, 0, TAG_B_ACC);
      tc_fence_after();
    }
    const int n_rows = kDKV ? p.n_k : p.n_q;
    auto store_acc = [&](uint32_t tcol, float mul, void* out, int64_t sb, int64_t sh, int64_t sn) {
      uint8_t* dst = static_cast<uint8_t*>(out) + 2 * (batch * sb + head_r * sh + static_cast<int64_t>(my_row) * sn);
      constexpr int kCols = kHeadDim / kBwdSplit;          // accumulator columns per thread: [hh kCols, (hh + 1) kCols)
      constexpr int kChunk = kCols < 32 ? kCols : 32;      // 32 or 16 columns per TMEM load
#pragma unroll
      for (int c2 = 0; c2 < kCols / kChunk; ++c2) {
        const int col0 = hh * kCols + c2 * kChunk;
        uint32_t v[kChunk];
        if (have) {
          if constexpr (kChunk == 32) tmem_ld32(tmem_base + lane_base + tcol + col0, v);
          else tmem_ld16(tmem_base + lane_base + tcol + col0, v);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < kChunk; ++i) v[i] = 0u;
        }
        if (my_row < n_rows) {
#pragma unroll
          for (int q4 = 0; q4 < kChunk / 8; ++q4) {
            const int col = col0 + q4 * 8;
            if (col < p.head_dim) {
              uint4 pk;
              pk.x = pack_16x2<kF16>(__uint_as_float(v[q4 * 8 + 0]) * mul, __uint_as_float(v[q4 * 8 + 1]) * mul);
              pk.y = pack_16x2<kF16>(__uint_as_float(v[q4 * 8 + 2]) * mul, __uint_as_float(v[q4 * 8 + 3]) * mul);
              pk.z = pack_16x2<kF16>(__uint_as_float(v[q4 * 8 + 4]) * mul, __uint_as_float(v[q4 * 8 + 5]) * mul);
              pk.w = pack_16x2<kF16>(__uint_as_float(v[q4 * 8 + 6]) * mul, __uint_as_float(v[q4 * 8 + 7]) * mul);
              *reinterpret_cast<uint4*>(dst + col * 2) = pk;
            }
          }
        }
      }
    };
    if (kDKV) {
      store_acc(T::kTmemAcc, 1.0f, p.out0, p.o0_sb, p.o0_sh, p.o0_sn);                    // dV
      store_acc(T::kTmemAcc + kHeadDim, p.scale, p.out1, p.o1_sb, p.o1_sh, p.o1_sn);      // dK
    } else {
      store_acc(T::kTmemAcc, p.scale, p.out0, p.o0_sb, p.o0_sh, p.o0_sn);                 // dQ
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdProdWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[row] = sum_c dO[row, c] * O[row, c] and LSE2[row] = LSE[row] * log2(e) into the padded workspace.  A row is read by `tpr`
// neighbouring threads (a power of two >= head_dim / 8), 16 bytes each, so a warp's loads cover whole contiguous rows.
template <typename T16>
__global__ void __launch_bounds__(256)
fa_bwd_prep_kernel(const T16* __restrict__ o, int64_t o_sb, int64_t o_sh, int64_t o_sn, const T16* __restrict__ d_o, int64_t g_sb,
                   int64_t g_sh, int64_t g_sn, const float* __restrict__ lse, float* __restrict__ l2, float* __restrict__ dsum,
                   int heads, int n_q, int n_q_pad, int d, int tpr, int64_t rows_pad) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t idx = tid / tpr;          // padded row
  const int piece = static_cast<int>(tid % tpr);
  const bool live = idx < rows_pad;       // (the shuffles below need every lane of the warp)
  const int64_t bh = live ? idx / n_q_pad : 0;
  const int row = live ? static_cast<int>(idx % n_q_pad) : n_q;
  float acc = 0.f;
  if (row < n_q && piece * 8 < d) {
    const int64_t b = bh / heads, h = bh % heads;
    const uint4 vo = *reinterpret_cast<const uint4*>(o + b * o_sb + h * o_sh + static_cast<int64_t>(row) * o_sn + piece * 8);
    const uint4 vg = *reinterpret_cast<const uint4*>(d_o + b * g_sb + h * g_sh + static_cast<int64_t>(row) * g_sn + piece * 8);
    const T16* po = reinterpret_cast<const T16*>(&vo);
    const T16* pg = reinterpret_cast<const T16*>(&vg);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(ld_as_float(po + i), ld_as_float(pg + i), acc);
  }
  for (int off = tpr >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (live && piece == 0) {
    float l = INFINITY;
    if (row < n_q) {
      const float x = lse[bh * n_q + row];
      l = (x == -INFINITY) ? INFINITY : x * 1.4426950408889634f;
    }
    l2[idx] = l;
    dsum[idx] = acc;
  }
}

}  // namespace fa
