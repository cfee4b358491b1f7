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
// Work items.  An item is 256 query rows of one (batch, head): two 128-row tiles in tile slots A and B that
// ping-pong on the tensor pipe — while the softmax warps of one slot work on S_j, the tensor pipe runs P*V and the
// next Q*K^T of the other.  The host sizes the item list as whole waves of 256-row items plus a remainder wave of
// 128-row items.  A 128-row item runs in "split-KV" mode: slot A attends the first half of the K/V tiles and slot B
// the second half (same Q rows), so the two halves ping-pong exactly like two Q tiles do, and the partial (O, m, l)
// pairs are merged by the log-sum-exp rule in the epilogue (slot A reads O_B straight from TMEM: both slots use the
// same TMEM lanes, different columns).  The serial chain of a tail item is therefore half as long.
//
// Persistent CTAs.  The grid is one CTA per SM; every CTA loops over items it takes from a global atomic counter
// (heaviest causal blocks first), so TMEM, barriers and tensor maps are set up once per SM, the TMA producer runs ahead
// into the next item's K/V (and Q, where SMEM has room for a second Q set) while the current item is still in its
// K/V loop or epilogue, and causal work stays balanced.  The item index travels from the producer to the other warps
// through a small SMEM queue (bar_wfull / bar_wempty).  With FwdParams::work_counter == nullptr items are taken with
// a static stride instead (one item per CTA when the grid is as large as the item list).
//
// Warp roles (320 threads): warps 0-3 softmax/epilogue of slot A, 4-7 of slot B (warp w reads TMEM lanes
// 32*(w%4)..+31), warp 8 lane 0 = TMA producer + item fetch (+ TMEM alloc/dealloc by the whole warp), warp 9 = MMA
// issuer (converged warp, one elected lane issues A then B in order).  `tcgen05.mma` issue back-pressures at the rate
// the tensor pipe retires (~64 cycles per 128x128x16 MMA) and every mbarrier wait + tcgen05 fence costs ~200 cycles
// even when already complete (measured: profiles/r01_trace_v1_c4_report.txt), so waits are batched.  Two
// independent issuer warps (one per slot) were measured and rejected: the slots fall into lock-step
// (profiles/r01_ab_issuer_modes_session9.log).
//
// TMEM columns (512 allocated): S_A [0,128)  S_B [128,256)  O_A [256,256+d)  O_B [256+d, 256+2d).
// P aliases S: the 16-bit paths (bf16, fp16) pack two values per column into S cols [0,64); tf32 overwrites S in place.
//
// Instances: tile rows of 128, 256 or 512 bytes (fp32 d = 32 / 64 / 128, 16-bit d = 64 / 128 / 256).  The 512-byte-row
// instances are "one-slot" (FwdTraits::kSlots == 1): one Q tile and a two-tile K/V ring fill SMEM, every item is a 128-row item
// on slot A, slot B's warps idle (and O_B's columns stay unused, which is what lets d = 256 fit).  Head dims below an
// instance run on it unchanged: the tensor maps carry the true head dim, TMA zero-fills the missing columns on load and
// clips them on store.
//
// SMEM (dynamic, 1024-B aligned): kQSets x (Q_A | Q_B) (Q_A only with 512-byte rows) | ring of NBUF K/V tiles | barriers | work queue | (m, l)
// exchange.  Every tile is DCHUNKS boxes of [128 rows x 128 bytes] in the SWIZZLE_128B layout that TMA writes and
// the UMMA descriptors read.  A slot's Q buffer doubles as the staging buffer of its O tile for the TMA store.
#pragma once
#include "ptx.cuh"

// bring-up / tuning switches (A/B-tested on hardware; the defaults are what ships)
#ifndef FA_OPT_SPLITP
#define FA_OPT_SPLITP 1   // deliver P to the MMA thread in two 64-key halves so P*V overlaps the second half of the exps
#endif
#ifndef FA_OPT_SPLITP_NARROW
#define FA_OPT_SPLITP_NARROW 1   // 0: the 128-byte-row instances (fp32 d <= 32, 16-bit d <= 64) hand P over in one piece: their P*V is
                                 // short, and the MMA warp — busy for the whole step at these head dims — saves a wait and an
                                 // elect block per slot and step
#endif
#ifndef FA_OPT_LDPIPE
#define FA_OPT_LDPIPE 0   // overlap the row-max pass with the remaining tcgen05.ld of the S row
#endif
#ifndef FA_OPT_F2
#define FA_OPT_F2 1       // packed FFMA2 / FADD2 for the scale-subtract and the row sum (bf16 instances only: in the tf32
                          // instances the per-element P truncation breaks register pairing and costs ~150 extra moves)
#endif
#ifndef FA_OPT_TF32_COMP
#define FA_OPT_TF32_COMP 1  // tf32 instances: instead of truncating every P element to tf32 before it is summed (one LOP3 per
                            // element, which also keeps the packed FFMA2/FADD2 forms from being used), P is computed as
                            // P*(1+eps) — eps = the mean relative truncation error of a tf32 operand, folded into the exp2
                            // argument for free — so that what the tensor core reads (P*(1+eps) truncated) is P on average, and
                            // the row sum of the untruncated values is divided by (1+eps) once per row in the epilogue
#endif
#ifndef FA_OPT_EARLY_S
#define FA_OPT_EARLY_S 0  // instances with spare TMEM columns (head dim <= 64): the second piece of P is written to its own
                          // columns instead of over S, so the next Q*K^T no longer has to wait for it — it is issued right
                          // after the first piece's P*V, while the softmax warps are still in the exps of the second piece.
                          // Correct (62 GPU tests) but measured 3-7% slower (profiles/r01_ab_early_s.log): with one in-order
                          // issuer the other slot's work queues behind this slot's second-piece wait
#endif
#ifndef FA_OPT_LATE_QFREE
#define FA_OPT_LATE_QFREE 1  // two-Q-set instances: wait for an item's O store (and release its staging buffer) after the first
                             // softmax step of the next item instead of inside the epilogue
#endif
#ifndef FA_OPT_RELEASE_IN_STEP
#define FA_OPT_RELEASE_IN_STEP 1  // K/V ring slots released from inside slot B's step (one elect block less per K/V step)
#endif
#ifndef FA_OPT_EARLY_HI
#define FA_OPT_EARLY_HI 0    // 16-bit instances: P (two values per column) only covers columns [0, 64) of S, and once the first piece
                             // of P has arrived every S column is in registers — so the upper half of the next Q*K^T (keys 64..127
                             // -> columns [64, 128)) is issued right behind the first piece's P*V instead of after the second's
#endif
#ifndef FA_OPT_ROT_S
#define FA_OPT_ROT_S 0       // 1: two-slot instances with head dim <= 64 rotate the S tiles of both slots through THREE TMEM buffers
                             // (3 * 128 + 2 * d <= 512 columns), so Q K^T of a slot's next step is issued a whole step early — behind
                             // the other slot's P V, into the buffer that P V has just freed — and drops out of the slot's serial
                             // chain softmax(j) -> P V(j) -> Q K^T(j+1) -> softmax(j+1).  Correct (161 GPU tests) and measured on B200
                             // (profiles/r02_ab_rotating_s.log): fp32 d=32 3% faster, fp32 d=64 18-20% SLOWER, bf16 d=64 7% slower.
                             // The chain is not what bounds these instances: their tensor work is instruction-rate bound (a
                             // tcgen05.mma takes >= 47 cycles however small N is, harness/mma_rate_probe.cu: P V at d <= 64 is 16
                             // such instructions per tile), the in-order issuer is busy for the whole step, and with S early both
                             // slots' softmaxes run at the same time and share the MUFU instead of alternating with the MMAs.
#endif
// K/V ring depth of the 128- and 256-byte-row instances (what fits next to the Q tiles: 2 x 2 x 16 KB + 8 x 16 KB, 2 x 32 KB + 5 x 32 KB).
// The MMA warp waits ~500 cycles per step for tiles that were requested 1 - 1.5 steps earlier (profiles/r02_ab_ring_depth.log).
#ifndef FA_NBUF_NARROW
#define FA_NBUF_NARROW 8
#endif
#ifndef FA_NBUF_MID
#define FA_NBUF_MID 5
#endif
#ifndef FA_OPT_ROLL_MMA
#define FA_OPT_ROLL_MMA 0    // 1: the k-step loops of the MMA warp stay rolled (smaller code for a warp that shares its instruction
                             // cache with the unrolled exp loops of the softmax warps)
#endif
#ifndef FA_OPT_ROLL_STEP
#define FA_OPT_ROLL_STEP 0   // 1: one copy of the per-slot step code, looped over the two slots
#endif
#if FA_OPT_ROLL_MMA
#define FA_MMA_UNROLL _Pragma("unroll 1")
#else
#define FA_MMA_UNROLL _Pragma("unroll")
#endif
#if FA_OPT_ROLL_STEP
#define FA_STEP_UNROLL _Pragma("unroll 1")
#else
#define FA_STEP_UNROLL _Pragma("unroll")
#endif
#ifndef FA_OPT_SPLIT_KEYS
#define FA_OPT_SPLIT_KEYS 64  // P is handed to the MMA warp in two pieces: keys [0, FA_OPT_SPLIT_KEYS) and the rest (64 or 96).  16-bit instances
#endif
#ifndef FA_OPT_SPLIT_KEYS_TF32
#define FA_OPT_SPLIT_KEYS_TF32 96   // tf32 instances.  By ncu launch durations (profiles/r02_fwd_ab_knobs_ncu.log): 96 is 1 % (C2) to 2.5 % (C3)
                                    // faster than 64 on the tf32 instances and 1.9 % slower on bf16 d = 128 (the event-timed A/B of round 1,
                                    // profiles/r01_ab_tf32_comp.log, could not resolve this)
#endif
// Of every FA_POLY_DEN_x consecutive element pairs of a P row, the first FA_POLY_NUM_x get exp2 from the FMA-pipe
// polynomial (packed FFMA2/FADD2) instead of MUFU.EX2: the softmax is MUFU-bound (16 ex2/clk/SM), the polynomial moves part
// of that load to the FMA pipe.  Measured on B200 (profiles/r01_ab_poly_persistent.log, r01_ab_tf32_comp.log):
// 1/4 is +7% (bf16 d128) / +9% (tf32 d64), 1/2 is slower than none.
#ifndef FA_POLY_NUM_BF16
#define FA_POLY_NUM_BF16 1
#endif
#ifndef FA_POLY_DEN_BF16
#define FA_POLY_DEN_BF16 4
#endif
#ifndef FA_POLY_NUM_TF32
#define FA_POLY_NUM_TF32 1
#endif
#ifndef FA_POLY_DEN_TF32
#define FA_POLY_DEN_TF32 4
#endif
// The 128-byte-row instances (fp32 d <= 32, 16-bit d <= 64) leave the tensor pipe mostly idle and run both slots' softmaxes
// side by side, so their MUFU / FMA balance is tuned separately.
#ifndef FA_POLY_NUM_NARROW
#define FA_POLY_NUM_NARROW 1
#endif
#ifndef FA_POLY_DEN_NARROW
#define FA_POLY_DEN_NARROW 4
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
  int tail_split;     // != 0: a 128-row item splits its K/V range over tile slots A and B (merged in the epilogue);
                      // 0: it runs slot A only
  int n_items;        // n_big + 2 * (number of 256-row blocks run as pairs of 128-row items)
  unsigned int* work_counter;   // [0] next item, [1] CTAs finished (the last one resets both); nullptr: static stride
  void* o_ptr;        // O in global memory with its element strides: only for rows that see no key (zeros, no staging buffer)
  int64_t o_sb, o_sh, o_sn;
  int o_row_bytes;    // head_dim * sizeof(output element), a multiple of 16
  // Accumulate mode (the steps of a ring forward after the first): an earlier partial result over other keys — O_acc fp32
  // [batch, heads, n_q, head_dim] contiguous, already normalised, and its LSE_acc [batch, heads, n_q] — is folded into this
  // launch's result by the log-sum-exp rule in the epilogue, before the O tile is staged:  lse = log(e^lse_acc + e^lse_new),
  // O = O_acc e^(lse_acc - lse) + O_new e^(lse_new - lse).  O may be O_acc itself (each tile is read by the CTA that
  // overwrites it, before it does) and `lse` may be LSE_acc.  nullptr: plain forward.
  const float* acc_o;
  const float* acc_lse;
  int head_dim;       // true head dim (<= the instance's): row pitch of O_acc
  // Split-KV across CTAs (launches with far fewer Q tiles than SMs and many K/V tiles — decode-like shapes): the K/V tiles of
  // every Q tile are cut into kv_splits runs of kv_chunk_tiles; split s is an item of its own that attends run s only and
  // writes its normalised partial O (fp32) and LSE into slice s of a workspace shaped [kv_splits * batch, heads, n_q, d]
  // (the O tensor map and `lse` describe that workspace); fa_combine_splits_kernel merges the slices.  1 = off.
  int kv_splits, kv_chunk_tiles;
  int kv_group;       // query heads per K/V head (grouped-query / multi-query attention): K/V head of query head h is h / kv_group; 1 = MHA
};
constexpr int kTraceSteps = 48;

constexpr int kBlockM = 128;          // rows per Q tile
constexpr int kBlockN = 128;          // keys per K/V tile
constexpr int kChunkBytes = 128 * 128;  // one TMA box: 128 rows x 128 bytes
constexpr int kNumThreads = 320;      // 8 softmax warps + TMA producer warp + MMA-issuer warp
constexpr int kBarMerge = 3;          // named barrier: slot B hands its partial (m, l) to slot A (split-KV tail items)
constexpr int kWorkQueue = 4;         // depth of the SMEM item queue (the producer is at most two items ahead)
constexpr float kRescaleThreshold = 8.0f;  // lazy rescale: keep a stale max while it is within 2^8
// tf32 truncation compensation (FA_OPT_TF32_COMP): an operand with a log-uniform mantissa loses on average
// eps = 2^-11 / (2 ln 2) of its value when the tensor core drops the low 13 mantissa bits
constexpr float kTf32CompEps = 3.5222e-4f;
constexpr float kTf32CompLog2 = 5.0806e-4f;          // log2(1 + eps), added to the exp2 argument
constexpr float kTf32CompInv = 1.0f / (1.0f + kTf32CompEps);

// kPrecise (fp32 inputs, head dim <= 64): "3xTF32".  kind::tf32 reads the top 19 bits of an fp32 operand, so a product of two
// fp32 values loses ~2^-11 per operand.  Every operand x is therefore used twice: as it is (the tensor core sees hi = trunc(x))
// and as lo = x - trunc(x) (exact in fp32, 11 more bits once truncated), and each contraction runs three MMAs into the same
// accumulator — hi*hi + lo*hi + hi*lo — which carries ~21 bits per operand, i.e. fp32-grade products.  A tile is then the raw
// chunks followed by the same number of lo chunks; TMA fills the raw half (plain FLOAT32 maps: no rounding on the way in) and
// the four warps of slot B — idle in a one-slot instance — write the lo half.  P: hi over S (in place), lo in slot B's S columns.
template <bool kTF32, int kHeadDim, bool kOutF32, bool kPrecise = false>
struct FwdTraits {
  static constexpr int kInSize = kTF32 ? 4 : 2;
  static constexpr int kOutSize = (kTF32 || kOutF32) ? 4 : 2;
  static constexpr int kDChunks = kHeadDim * kInSize / 128;    // boxes of Q/K/V data per tile (what TMA loads)
  static constexpr int kTileChunks = kPrecise ? 2 * kDChunks : kDChunks;   // + the lo copies
  static constexpr int kOChunks = kHeadDim * kOutSize / 128;   // boxes per O tile
  static constexpr int kElemsPerChunk = 128 / kInSize;
  static constexpr int kOutElemsPerChunk = 128 / kOutSize;
  static constexpr int kLoadBytes = kDChunks * kChunkBytes;    // bytes TMA delivers per tile
  static constexpr int kLoOffset = kDChunks * kChunkBytes;     // lo copy of a tile, relative to the tile (kPrecise)
  static constexpr int kTileBytes = kTileChunks * kChunkBytes;
  // Tile slots with their own Q buffer.  512-byte rows (fp32 d = 128, 16-bit d = 256, and fp32 d = 64 with its lo copy) leave
  // SMEM for one Q tile and a two-tile K/V ring only: those instances run every item as a 128-row item on slot A (the host
  // never builds 256-row or split-KV items for them) and slot B's warps idle.  K_(j+1) then streams in under the softmax of step j and
  // V_(j+1) under Q K^T(j+1) and that softmax — the tensor pipe has nothing else to do for a lone Q tile anyway.
  static constexpr int kSlots = kTileChunks <= 2 ? 2 : 1;
  static constexpr int kNBuf = kTileChunks == 1 ? FA_NBUF_NARROW : (kTileChunks == 2 ? FA_NBUF_MID : 2);   // K/V ring depth (tiles)
  static constexpr int kUmmaK = 32 / kInSize;                  // K per tcgen05.mma: 8 (tf32) / 16 (bf16)
  static constexpr int kQSets = kTileChunks == 1 ? 2 : 1;      // Q double-buffered across items where SMEM allows
  static constexpr bool kComp = kTF32 && !kPrecise && (FA_OPT_TF32_COMP != 0);   // tf32 truncation compensated instead of reproduced
  static constexpr bool kPacked = (FA_OPT_F2 != 0) && (!kTF32 || kComp || kPrecise);   // FFMA2 / FADD2 forms in the exp loop
  static constexpr int kSplitKeys = kTF32 ? FA_OPT_SPLIT_KEYS_TF32 : FA_OPT_SPLIT_KEYS;
  static_assert(kSplitKeys == 64 || kSplitKeys == 96, "P split point");
  // polynomial exp2 on kPolyNum of every kPolyDen element pairs (packed path only; never in a precise instance: its 7.5e-5
  // relative error is what that mode exists to avoid)
  static constexpr int kPolyNum = kPrecise ? 0 : (kDChunks == 1 ? FA_POLY_NUM_NARROW : (kTF32 ? FA_POLY_NUM_TF32 : FA_POLY_NUM_BF16));
  static constexpr int kPolyDen = kDChunks == 1 ? FA_POLY_DEN_NARROW : (kTF32 ? FA_POLY_DEN_TF32 : FA_POLY_DEN_BF16);
  static constexpr bool kRotS = (FA_OPT_ROT_S != 0) && kSlots == 2 && (FA_OPT_EARLY_S == 0) && (FA_OPT_EARLY_HI == 0) &&
                                (3 * kBlockN + 2 * kHeadDim <= 512);
  static constexpr int kSBufs = kRotS ? 3 : 2;   // S tile n = 2 * step + slot of an item lives in buffer n % kSBufs
  static constexpr int kSmemData = (kSlots * kQSets + kNBuf) * kTileBytes;
  static constexpr int kNumBarriers = 4 * kQSets /*q full, q free*/ + 2 * kNBuf + 2 /*s_full*/ + 4 /*p_full halves*/ +
                                      2 /*o_final*/ + 2 /*o_free*/ + 2 * kWorkQueue + 2 /*pv1 done*/ +
                                      (kPrecise ? kNBuf + 2 : 0) /*lo copy of a K/V tile / of the Q tile written*/ +
                                      (kRotS ? 6 : 0) /*s_full and p_full of the odd steps*/;
  static constexpr int kSmemBytes = kSmemData + kNumBarriers * 8 + 16 /*tmem ptr*/ + kWorkQueue * 4 + 2 * kBlockM * 4 /*m, l*/ +
                                    1024 /*alignment slack*/;
  static constexpr int kTmemS = 0;        // + 128*t (kRotS: + 128 * (tile number % 3))
  static constexpr int kTmemO = kRotS ? 384 : 256;      // + kHeadDim*t
  static constexpr int kTmemPLo = kBlockN;   // kPrecise: lo part of P, in the S columns of the (unused) slot B
  static constexpr int kP1Cols = (kBlockN - kSplitKeys) * kInSize / 4;   // TMEM columns of the second piece of P
  static constexpr int kTmemP1 = 256 + 2 * kHeadDim;                      // + kP1Cols*t (kEarlyS only)
  static constexpr bool kSplitP = (FA_OPT_SPLITP != 0) && (kDChunks > 1 || FA_OPT_SPLITP_NARROW != 0);   // P delivered in two pieces
  static constexpr bool kEarlyHi = (FA_OPT_EARLY_HI != 0) && kSplitP && !kTF32 && (FA_OPT_EARLY_S == 0);
  static constexpr bool kEarlyS = (FA_OPT_EARLY_S != 0) && !kPrecise && kSplitP && (256 + 2 * kHeadDim + 2 * kP1Cols <= 512);
  static_assert(kDChunks == 1 || kDChunks == 2 || kDChunks == 4, "tile row must be 128, 256 or 512 bytes");
  static_assert(!kPrecise || (kTF32 && kSlots == 1 && !kOutF32), "precise instances: fp32 operands, one slot (slot B's warps write the lo copies)");
  static_assert(kTmemO + kSlots * kHeadDim <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 227 * 1024, "SMEM budget");
  // SMEM tile index of Q buffer qb = set * 2 + slot (the barrier index): one-slot instances have no tile for slot B
  __host__ __device__ static constexpr int q_tile(int qb) { return kSlots == 2 ? qb : (qb >> 1); }
};

// watchdog tags
enum : uint32_t {
  TAG_Q_FULL = 1, TAG_KV_FULL = 2, TAG_KV_EMPTY = 3, TAG_S_FULL = 4, TAG_P_FULL = 5, TAG_O_FINAL = 6, TAG_Q_FREE = 7,
  TAG_O_FREE = 8, TAG_W_FULL = 9, TAG_W_EMPTY = 10, TAG_PV1 = 11, TAG_CONV = 12
};

// What one work item is, derived from its index by every warp role on its own.
struct Item {
  int head, batch, row0;   // first query row of slot A
  int row1;                // first query row of slot B (== row0 in split mode)
  bool single, split;
  int n0, n1;              // K/V tiles of slot A / slot B (scalars: a runtime-indexed array would live in local memory)
  int kv_first1;           // first K/V tile of slot B (split mode), else 0
  int kv_base;             // first K/V tile of the item (cross-CTA split-KV: run `kv_split` of the key sequence), else 0
  int obatch;              // batch coordinate of the item's output rows: batch, or kv_split * batch_count + batch in the workspace
  int n_max;
  FA_DEVINL int n(int t) const { return t == 0 ? n0 : n1; }
  // Slot t gets a Q tile of its own in this item (in split mode slot B reads slot A's).  The ownership protocol of a Q buffer
  // (bar_q: tile landed; bar_qfree: the O tile staged there has been read out by its TMA store) has exactly one phase per
  // item for which this holds, on the producer's, the MMA warp's and the epilogue's side alike.
  FA_DEVINL bool loads_q(int t) const { return n(t) > 0 && !(split && t == 1); }
};
template <bool kCausal>
FA_DEVINL Item decode_item(const FwdParams& p, int bid) {
  Item it;
  it.single = bid >= p.n_big;
  int half = 0;
  if (it.single) {
    const int k = bid - p.n_big;
    half = k & 1;
    bid = p.n_big + (k >> 1);
  }
  int m_blk = bid % p.num_m_blocks;   // m fastest so neighbours share K/V in L2
  bid /= p.num_m_blocks;
  int kv_split = 0;
  if (p.kv_splits > 1) {
    kv_split = bid % p.kv_splits;
    bid /= p.kv_splits;
  }
  it.head = bid % p.heads;
  it.batch = bid / p.heads;
  it.kv_base = kv_split * p.kv_chunk_tiles;
  it.obatch = kv_split * p.batch + it.batch;
  if (kCausal) m_blk = p.num_m_blocks - 1 - m_blk;  // heaviest blocks first
  it.row0 = m_blk * (2 * kBlockM) + half * kBlockM;
  it.split = it.single && p.tail_split != 0;   // slots A and B = two halves of the K/V range of ONE Q tile
  const int n_kv_total = (p.n_k + kBlockN - 1) / kBlockN;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = it.row0 + t * kBlockM;
    int n = n_kv_total;
    if (kCausal) {
      const int last_key = r0 + kBlockM - 1 + p.causal_offset;
      n = last_key < 0 ? 0 : min(n_kv_total, last_key / kBlockN + 1);
    }
    if (p.kv_splits > 1) n = max(0, min(n - it.kv_base, p.kv_chunk_tiles));   // this run's share of the visible tiles
    if (r0 >= p.n_q || (it.single && t == 1)) n = 0;
    if (t == 0) it.n0 = n; else it.n1 = n;
  }
  it.kv_first1 = 0;
  if (it.split) {
    const int n = it.n0;
    it.n0 = (n + 1) >> 1;
    it.n1 = n - it.n0;
    it.kv_first1 = it.n0;
  }
  it.n_max = max(it.n0, it.n1);
  it.row1 = it.split ? it.row0 : it.row0 + kBlockM;
  return it;
}

// Every MMA group of an item in issue order, for instances whose S tiles rotate through kSBufs TMEM buffers (kRotS).
// Tile n = 2 * step + slot.  f(true, t, j): P_t(j) V -> O_t;  f(false, t, j): S_t(j) = Q_t K^T into buffer n % kSBufs.
// S of tile n + kSBufs goes into the buffer that P V of tile n has just finished reading (the tensor pipe runs in issue
// order), so every slot's S is a full step ahead of its softmax.  The TMA producer walks the same sequence to load the
// K/V tiles in the order of their first use.
template <int kSBufs, typename F>
FA_DEVINL void for_each_mma(const Item& w, F&& f) {
  const int n_tiles = 2 * w.n_max;
#pragma unroll 1
  for (int n = -kSBufs; n < n_tiles; ++n) {
    if (n >= 0) {
      const int t = n & 1, j = n >> 1;
      if (j < w.n(t)) f(true, t, j);
    }
    const int m = n + kSBufs;
    if (m < n_tiles) {
      const int t = m & 1, j = m >> 1;
      if (j < w.n(t)) f(false, t, j);
    }
  }
}
// In a 256-row item both slots read the same K/V tiles: slot A's use of tile j is the first (slot B's when A has no step j),
// slot B's the last (slot A's when B has none).  In a split-KV item every slot has tiles of its own.
FA_DEVINL bool kv_first_use(const Item& w, int t, int j) { return w.split || t == 0 || j >= w.n0; }
FA_DEVINL bool kv_last_use(const Item& w, int t, int j) { return w.split || t == 1 || j >= w.n1; }

// kF16 (16-bit instances only): the operands are IEEE fp16 instead of bf16 — same kind::f16 instruction, operand format 0
// instead of 1; P <= 2^kRescaleThreshold by construction (lazy rescale), far inside the fp16 range.
template <bool kTF32, int kHeadDim, bool kCausal, bool kOutF32, bool kF16 = false, bool kPrecise = false>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                    const FwdParams p) {
  using T = FwdTraits<kTF32, kHeadDim, kOutF32, kPrecise>;
  constexpr int kQS = T::kQSets;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                              // [kQS][kSlots] tiles
  const uint32_t sKV = smem_base + T::kSlots * kQS * T::kTileBytes;   // kNBuf tiles
  const uint32_t sBar = smem_base + T::kSmemData;
  const uint32_t bar_q = sBar;                                // [kQS][2]  Q tile landed
  const uint32_t bar_qfree = bar_q + 16 * kQS;                // [kQS][2]  Q buffer (= O staging) reusable
  const uint32_t bar_full = bar_qfree + 16 * kQS;             // [kNBuf]
  const uint32_t bar_empty = bar_full + 8 * T::kNBuf;         // [kNBuf]
  const uint32_t bar_s = bar_empty + 8 * T::kNBuf;            // [2]
  const uint32_t bar_p = bar_s + 16;                          // [slot][half] = [4]
  const uint32_t bar_o = bar_p + 32;                          // [2]  O_t final
  const uint32_t bar_ofree = bar_o + 16;                      // [2]  slot t's epilogue has read its accumulator(s) out of TMEM
  const uint32_t bar_wfull = bar_ofree + 16;                  // [kWorkQueue]
  const uint32_t bar_wempty = bar_wfull + 8 * kWorkQueue;     // [kWorkQueue]
  const uint32_t bar_pv1 = bar_wempty + 8 * kWorkQueue;       // [2]  second-piece P*V of slot t's latest step has completed
  const uint32_t bar_conv = bar_pv1 + 16;                     // [kNBuf]  kPrecise: lo copy of the K/V tile in ring slot i written
  const uint32_t bar_qconv = bar_conv + 8 * T::kNBuf;         // [2]      kPrecise: lo copy of the Q tile written ([1] unused)
  const uint32_t bar_s_odd = bar_pv1 + 16;                    // [2]      kRotS: S of the odd steps (a slot's S runs a step ahead of its
  const uint32_t bar_p_odd = bar_s_odd + 16;                  // [slot][half]    softmax, so consecutive steps need barriers of their own)
  const uint32_t s_tmem_ptr = kPrecise ? bar_qconv + 16 : (T::kRotS ? bar_p_odd + 32 : bar_pv1 + 16);   // 16 bytes
  const uint32_t s_work = s_tmem_ptr + 16;                    // [kWorkQueue] item indices (-1 = no more work)
  const uint32_t s_ml = s_work + 4 * kWorkQueue;              // m[128], l[128] of slot B (split-KV merge)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FA_TRACE_MISC(0, 0);
  if (threadIdx.x == 128) FA_TRACE_MISC(1, 0);
  if (warp == 9) FA_TRACE_MISC(2, 0);

  // ---- one-time setup ----
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2 * kQS; ++i) {
      mbar_init(bar_q + 8 * i, 1);
      mbar_init(bar_qfree + 8 * i, 1);
    }
    for (int i = 0; i < T::kNBuf; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_s + 8 * t, 1);
      mbar_init(bar_p + 16 * t, 128);
      mbar_init(bar_p + 16 * t + 8, 128);
      mbar_init(bar_o + 8 * t, 1);
      mbar_init(bar_pv1 + 8 * t, 1);
    }
    mbar_init(bar_ofree, 4);              // one arrival per softmax warp of the slot, only in items where it read TMEM
    mbar_init(bar_ofree + 8, 4);
    for (int i = 0; i < kWorkQueue; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 9);   // MMA warp + 8 softmax warps
    }
    if constexpr (T::kRotS) {
      for (int t = 0; t < 2; ++t) {
        mbar_init(bar_s_odd + 8 * t, 1);
        mbar_init(bar_p_odd + 16 * t, 128);
        mbar_init(bar_p_odd + 16 * t + 8, 128);
      }
    }
    if constexpr (kPrecise) {
      for (int i = 0; i < T::kNBuf; ++i) mbar_init(bar_conv + 8 * i, 4);   // one arrival per warp of slot B
      mbar_init(bar_qconv, 4);
      mbar_init(bar_qconv + 8, 4);
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

  // s_full / p_full of step g of a slot (g counts the slot's K/V steps over all items).  kRotS: even and odd steps have
  // barriers of their own — S_t(g+1) may complete before the softmax has waited for S_t(g), and P_t(g+1) may be delivered
  // before the MMA warp has waited for P_t(g); a single barrier would then be two phases ahead of its waiter.
  auto bar_s_at = [&](int t, int g) { return ((T::kRotS && (g & 1)) ? bar_s_odd : bar_s) + 8 * t; };
  auto bar_p_at = [&](int t, int h, int g) { return ((T::kRotS && (g & 1)) ? bar_p_odd : bar_p) + 16 * t + 8 * h; };
  auto par_at = [&](int g) { return static_cast<uint32_t>(T::kRotS ? (g >> 1) & 1 : g & 1); };
  // consumer side of the item queue: every lane of the warp reads the item, one lane releases the queue slot
  auto next_item = [&](int seq) -> int {
    if (seq == 0) return static_cast<int>(blockIdx.x);
    const int ws = seq % kWorkQueue;
    // queue slot 0 is first used by seq = kWorkQueue (seq 0 bypasses the queue), the others by seq = ws
    mbar_wait(bar_wfull + 8 * ws, (seq / kWorkQueue - (ws == 0 ? 1 : 0)) & 1, TAG_W_FULL);
    const int item = static_cast<int>(ld_shared_b32(s_work + 4 * ws));
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_wempty + 8 * ws);
    return item;
  };

  if (warp == 8) {
    // =========================== item fetch + TMA producer ===========================
    if (lane == 0) {
      int ring = 0;   // running index into the K/V ring; the MMA issuer consumes tiles in exactly this order
      uint32_t qf_used = 0, qf_par = 0;   // per Q buffer: a tile has been loaded before / parity of the next bar_qfree phase
      for (int seq = 0;; ++seq) {
        // the first item of a CTA is its block index (no round trip to the counter on the critical start-up path, and
        // no queue: every role knows it); later items come from the counter, or from a static stride without one
        int item = static_cast<int>(blockIdx.x);
        if (seq > 0) {
          if (p.work_counter != nullptr) item = static_cast<int>(gridDim.x + atomicAdd(p.work_counter, 1u));
          else item = static_cast<int>(blockIdx.x) + seq * static_cast<int>(gridDim.x);
          if (item >= p.n_items) item = -1;
          const int ws = seq % kWorkQueue;
          const int use = seq / kWorkQueue - (ws == 0 ? 1 : 0);   // how often this queue slot has been used before
          if (use > 0) mbar_wait(bar_wempty + 8 * ws, (use - 1) & 1, TAG_W_EMPTY);
          st_shared_b32(s_work + 4 * ws, static_cast<uint32_t>(item));
          mbar_arrive(bar_wfull + 8 * ws);
        }
        if (item < 0) break;
        const Item w = decode_item<kCausal>(p, item);
        if (w.n_max == 0) continue;
        const int set = seq % kQS;
        auto load_kv = [&](const CUtensorMap* tm, int kv_tile) {
          const int buf = ring % T::kNBuf;
          const int round = ring / T::kNBuf;
          ++ring;
          if (round > 0) mbar_wait(bar_empty + 8 * buf, (round - 1) & 1, TAG_KV_EMPTY);
#if FA_TRACE
          if (seq == 0) FA_TRACE_AT(tm == &tm_k ? 2 : 3, kv_tile, 4);   // ring slot free -> TMA issued (first item: tile == step)
#endif
          mbar_arrive_expect_tx(bar_full + 8 * buf, T::kLoadBytes);
#pragma unroll
          for (int c = 0; c < T::kDChunks; ++c)
            tma_load_4d(sKV + buf * T::kTileBytes + c * kChunkBytes, tm, bar_full + 8 * buf, c * T::kElemsPerChunk,
                        (w.kv_base + kv_tile) * kBlockN, w.head / p.kv_group, w.batch);
        };
        // Q: one tile per slot; in split mode both slots read the same Q tile from slot A's buffer.  The buffer is
        // free once the epilogue that staged its O tile there (kQS items ago) has been read out by the TMA store.
        auto load_q = [&]() {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (w.loads_q(t)) {
              const int qb = set * 2 + t;
              // the previous tile in this buffer (if any) must have left as an O tile: one bar_qfree phase per load, counted
              // here and not derived from the item number — items and slots without a Q tile have no phase (a barrier that
              // runs a phase ahead of or behind its waiter makes the parity test alias)
              if (qf_used & (1u << qb)) {
                mbar_wait(bar_qfree + 8 * qb, (qf_par >> qb) & 1u, TAG_Q_FREE);
                qf_par ^= 1u << qb;
              }
              qf_used |= 1u << qb;
              mbar_arrive_expect_tx(bar_q + 8 * qb, T::kLoadBytes);
#pragma unroll
              for (int c = 0; c < T::kDChunks; ++c)
                tma_load_4d(sQ + T::q_tile(qb) * T::kTileBytes + c * kChunkBytes, &tm_q, bar_q + 8 * qb, c * T::kElemsPerChunk,
                            w.row0 + t * kBlockM, w.head, w.batch);
            }
          }
        };
        // The first kNBuf K/V tiles only need ring slots that the previous item releases on its own, so they are
        // issued before the Q tiles (whose buffer may still hold the previous item's O on its way out).
        int issued = 0;
        bool q_done = false;
        auto kv = [&](const CUtensorMap* tm, int kv_tile) {
          if (!q_done && issued == T::kNBuf) {
            load_q();
            q_done = true;
          }
          load_kv(tm, kv_tile);
          ++issued;
        };
        // With two Q sets the buffer is free long before (its O left two items ago) and on the very first item there is
        // nothing to wait for: Q first, it is needed first.  With one set the previous item's O is still on its way out.
        if (seq == 0 || kQS > 1) {
          load_q();
          q_done = true;
        }
        if constexpr (T::kRotS) {
          // tiles in the order of their first use by the MMA warp: K_0, K_1, V_0, K_2, V_1, ... (K runs a step ahead)
          for_each_mma<T::kSBufs>(w, [&](bool pv, int t, int j) {
            if (kv_first_use(w, t, j)) kv(pv ? &tm_v : &tm_k, (t == 0 ? 0 : w.kv_first1) + j);
          });
        } else if (!w.split) {
          for (int j = 0; j < w.n_max; ++j) {   // K_0, V_0, K_1, V_1, ... shared by both Q tiles
            kv(&tm_k, j);
            kv(&tm_v, j);
          }
        } else {
          // K_A0, K_B0, then per step: V_A(j), K_A(j+1), V_B(j), K_B(j+1)
          const int nA = w.n0, nB = w.n1;
          kv(&tm_k, 0);
          if (nB > 0) kv(&tm_k, nA);
          for (int j = 0; j < nA; ++j) {
            kv(&tm_v, j);
            if (j + 1 < nA) kv(&tm_k, j + 1);
            if (j < nB) {
              kv(&tm_v, nA + j);
              if (j + 1 < nB) kv(&tm_k, nA + j + 1);
            }
          }
        }
        if (!q_done) load_q();
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    // The whole warp follows the (warp-uniform) control flow and waits on the mbarriers; one elected lane issues.
    constexpr uint32_t kFmt = kTF32 ? 2u : (kF16 ? 0u : 1u);   // tf32 / fp16 / bf16 operands
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
    constexpr int kKStepsSplit = T::kSplitKeys / T::kUmmaK;   // k-steps covered by the first piece of P
    // Descriptors are built once per operand tile; a k-step only adds to the 14-bit start-address field
    // ((bytes >> 4); SMEM addresses are < 2^18, so the field never carries).
    auto issue_s = [&](int t, int qbuf, int buf) {
      const uint64_t qd = sdesc_at(hi_kmajor, sQ + T::q_tile(qbuf) * T::kTileBytes);
      const uint64_t kd = sdesc_at(hi_kmajor, sKV + buf * T::kTileBytes);
      const uint32_t d = tmem_base + T::kTmemS + t * kBlockN;
      FA_MMA_UNROLL
      for (int kk = 0; kk < kKStepsS; ++kk) {
        const uint32_t off16 = ((kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4;
        mma_ss<kTF32>(d, qd + off16, kd + off16, idesc_s, kk > 0 ? 1u : 0u);
      }
      if constexpr (kPrecise) {   // + lo(Q) K^T + Q lo(K)^T
        constexpr uint32_t lo16 = T::kLoOffset >> 4;
        FA_MMA_UNROLL
        for (int kk = 0; kk < kKStepsS; ++kk) {
          const uint32_t off16 = ((kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4;
          mma_ss<kTF32>(d, qd + lo16 + off16, kd + off16, idesc_s, 1u);
        }
        FA_MMA_UNROLL
        for (int kk = 0; kk < kKStepsS; ++kk) {
          const uint32_t off16 = ((kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4;
          mma_ss<kTF32>(d, qd + off16, kd + lo16 + off16, idesc_s, 1u);
        }
      }
    };
    // one 64-key half of S_t = Q_t K^T: keys [64*half, 64*half + 64) of the tile -> S columns [64*half, +64) (kEarlyHi).
    // K is K-major: key row r of a 128-byte column chunk sits at r * 128 bytes (swizzled inside 1024-byte groups of 8 rows)
    constexpr uint32_t idesc_s_half = make_idesc(kFmt, 0, kBlockM, kBlockN / 2);
    auto issue_s_half = [&](int t, int qbuf, int buf, int half) {
      const uint64_t qd = sdesc_at(hi_kmajor, sQ + T::q_tile(qbuf) * T::kTileBytes);
      const uint64_t kd = sdesc_at(hi_kmajor, sKV + buf * T::kTileBytes + half * (kBlockN / 2) * 128);
      const uint32_t d = tmem_base + T::kTmemS + t * kBlockN + half * (kBlockN / 2);
      FA_MMA_UNROLL
      for (int kk = 0; kk < kKStepsS; ++kk) {
        const uint32_t off16 = ((kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4;
        mma_ss<kTF32>(d, qd + off16, kd + off16, idesc_s_half, kk > 0 ? 1u : 0u);
      }
    };
    // P*V for k-steps [ks0, ks1) of the 128-key tile; A = P_t read from TMEM (it aliases S_t)
    auto issue_pv = [&](int t, int buf, bool accumulate, int ks0, int ks1) {
      const uint64_t vd = sdesc_at(hi_mnmajor, sKV + buf * T::kTileBytes);
      const uint32_t d = tmem_base + T::kTmemO + t * kHeadDim;
      // P aliases S; with kEarlyS the second piece (k-steps >= kKStepsSplit) has its own columns
      const bool own = T::kEarlyS && ks0 >= kKStepsSplit;
      const uint32_t a = own ? tmem_base + T::kTmemP1 + t * T::kP1Cols - kKStepsSplit * 8 : tmem_base + T::kTmemS + t * kBlockN;
      FA_MMA_UNROLL
      for (int ks = ks0; ks < ks1; ++ks) {
        mma_ts<kTF32>(d, a + ks * 8, vd + static_cast<uint32_t>(ks * (T::kUmmaK * 128 / 16)), idesc_pv,
                      (accumulate || ks > 0) ? 1u : 0u);
      }
      if constexpr (kPrecise) {   // + lo(P) V + P lo(V)
        constexpr uint32_t lo16 = T::kLoOffset >> 4;
        const uint32_t a_lo = tmem_base + T::kTmemS + T::kTmemPLo;
        FA_MMA_UNROLL
        for (int ks = ks0; ks < ks1; ++ks)
          mma_ts<kTF32>(d, a_lo + ks * 8, vd + static_cast<uint32_t>(ks * (T::kUmmaK * 128 / 16)), idesc_pv, 1u);
        FA_MMA_UNROLL
        for (int ks = ks0; ks < ks1; ++ks)
          mma_ts<kTF32>(d, a + ks * 8, vd + lo16 + static_cast<uint32_t>(ks * (T::kUmmaK * 128 / 16)), idesc_pv, 1u);
      }
    };
    // the same two with explicit TMEM addresses (kRotS: the S buffer changes from step to step)
    auto issue_s_at = [&](uint32_t d, int qbuf, int buf) {
      const uint64_t qd = sdesc_at(hi_kmajor, sQ + T::q_tile(qbuf) * T::kTileBytes);
      const uint64_t kd = sdesc_at(hi_kmajor, sKV + buf * T::kTileBytes);
      FA_MMA_UNROLL
      for (int kk = 0; kk < kKStepsS; ++kk) {
        const uint32_t off16 = ((kk >> 2) * kChunkBytes + (kk & 3) * 32) >> 4;
        mma_ss<kTF32>(d, qd + off16, kd + off16, idesc_s, kk > 0 ? 1u : 0u);
      }
    };
    auto issue_pv_at = [&](uint32_t d, uint32_t a, int buf, bool accumulate, int ks0, int ks1) {
      const uint64_t vd = sdesc_at(hi_mnmajor, sKV + buf * T::kTileBytes);
      FA_MMA_UNROLL
      for (int ks = ks0; ks < ks1; ++ks) {
        mma_ts<kTF32>(d, a + ks * 8, vd + static_cast<uint32_t>(ks * (T::kUmmaK * 128 / 16)), idesc_pv,
                      (accumulate || ks > 0) ? 1u : 0u);
      }
    };
    int gs_a = 0, gs_b = 0, gp_a = 0, gp_b = 0;   // kRotS: S tiles issued / P V groups issued per slot, over all items
    int ring = 0;                 // K/V ring index, same sequence as the producer's
    uint32_t p_par = 0;           // bit t: parity of the next bar_p[t] phase (one phase per K/V step of slot t, over all items)
    uint32_t q_par = 0;           // bit qb: parity of the next bar_q[qb] phase (one phase per Q tile loaded into buffer qb)
    uint32_t of_par = 0;          // bit t: parity of the next bar_ofree[t] phase
    uint32_t of_pend = 0;         // bit t: an epilogue of the previous item is (or will be) reading TMEM and arrives on bar_ofree[t]
    // The first P*V of an item overwrites O_t: every epilogue of the previous item that reads TMEM must be done with it.
    // Arrivals are tied to work (a slot that had none does not arrive), so a slot can never arrive twice in one phase.
    auto wait_ofree = [&]() {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (of_pend & (1u << t)) {
          mbar_wait(bar_ofree + 8 * t, (of_par >> t) & 1u, TAG_O_FREE);
          of_par ^= 1u << t;
        }
      }
      if (of_pend) tc_fence_after();
      of_pend = 0;
    };
#if FA_TRACE
    int steps_a = 0, steps_b = 0;
#endif
    // a K/V tile is usable once TMA has delivered it — in a precise instance, once its lo copy has been written next to it
    auto wait_full = [&](int i) {
      mbar_wait((kPrecise ? bar_conv : bar_full) + 8 * (i % T::kNBuf), (i / T::kNBuf) & 1, kPrecise ? TAG_CONV : TAG_KV_FULL);
    };
    // P_t(j) V -> O_t in two 64-key halves as the softmax warps deliver them, then (unless this was the slot's last
    // K/V tile of the item) S_t(j+1); the tensor pipe executes in issue order, so S_t(j+1) may overwrite the columns
    // P_t(j) aliased.
    auto step = [&](int t, int j, bool last, int qbuf, int vbuf, int kbuf, bool release) {
      const uint32_t par = (p_par >> t) & 1u;
      p_par ^= 1u << t;
#if FA_TRACE
      const int g = t == 0 ? steps_a++ : steps_b++;
#endif
      if constexpr (T::kEarlyS) {
        FA_TRACE_AT(2 + t, g, 0);
        mbar_wait(bar_p + 16 * t, par, TAG_P_FULL);            // first piece of P is in TMEM (over S), all of S_t(j) is in registers
        tc_fence_after();
        FA_TRACE_AT(2 + t, g, 1);
        if (elect_one_sync()) {
          issue_pv(t, vbuf, j > 0, 0, kKStepsSplit);
          if (!last) {
            issue_s(t, qbuf, kbuf);                            // in order after the P*V that read the aliased columns
            tc_commit(bar_s + 8 * t);
            if (release) tc_commit(bar_empty + 8 * kbuf);
          }
        }
        __syncwarp();
        FA_TRACE_AT(2 + t, g, 2);
        mbar_wait(bar_p + 16 * t + 8, par, TAG_P_FULL);        // second piece (own columns)
        tc_fence_after();
        FA_TRACE_AT(2 + t, g, 3);
        if (elect_one_sync()) {
          issue_pv(t, vbuf, true, kKStepsSplit, kKStepsPV);
          tc_commit(bar_pv1 + 8 * t);                          // the piece's columns are free again; O_t holds all of step j
          if (release) tc_commit(bar_empty + 8 * vbuf);
          if (last) tc_commit(bar_o + 8 * t);
        }
      } else {
      if constexpr (T::kSplitP) {
        FA_TRACE_AT(2 + t, g, 0);
        mbar_wait(bar_p + 16 * t, par, TAG_P_FULL);            // keys [0, 64) of P are in TMEM
        tc_fence_after();
        FA_TRACE_AT(2 + t, g, 1);
        if (elect_one_sync()) {
          issue_pv(t, vbuf, j > 0, 0, kKStepsSplit);
          if (T::kEarlyHi && !last) issue_s_half(t, qbuf, kbuf, 1);   // every column of S_t(j) is in registers by now
        }
        __syncwarp();
        FA_TRACE_AT(2 + t, g, 2);
      }
      mbar_wait(bar_p + 16 * t + 8, par, TAG_P_FULL);        // keys [64, 128)
      tc_fence_after();
      FA_TRACE_AT(2 + t, g, 3);
      if (elect_one_sync()) {
        issue_pv(t, vbuf, j > 0, T::kSplitP ? kKStepsSplit : 0, kKStepsPV);
        if (release) tc_commit(bar_empty + 8 * vbuf);
        if (last) {
          tc_commit(bar_o + 8 * t);
        } else {
          if constexpr (T::kEarlyHi) issue_s_half(t, qbuf, kbuf, 0);   // columns [0, 64): where P_t(j) was
          else issue_s(t, qbuf, kbuf);
          tc_commit(bar_s + 8 * t);
          if (release) tc_commit(bar_empty + 8 * kbuf);
        }
      }
      }
      __syncwarp();
      FA_TRACE_AT(2 + t, g, 6);
    };

    for (int seq = 0;; ++seq) {
      const int item = next_item(seq);
      if (item < 0) break;
      const Item w = decode_item<kCausal>(p, item);
      if (w.n_max == 0) continue;
      const int set = seq % kQS;
      // Q tiles of this item
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (w.loads_q(t)) {
          const int qb = set * 2 + t;
          mbar_wait((kPrecise ? bar_qconv : bar_q) + 8 * qb, (q_par >> qb) & 1u, kPrecise ? TAG_CONV : TAG_Q_FULL);
          q_par ^= 1u << qb;
        }
      }
      if (seq == 0) FA_TRACE_MISC(2, 1);
      if constexpr (T::kRotS) {
        // ---- S tiles rotate through three buffers (see for_each_mma, which the producer walks for the same tile order).
        // Iteration n: P V of tile n = (slot n & 1, step n >> 1), then — in the same elect block, so the warp's work per step
        // is what it is without the rotation — S of tile n + 3 into the buffer that P V has just finished with.
        bool ofree_done = false;
        int k_sh = 0, v_sh = 0;   // ring index of the K / V tile the two slots share (set by its first user)
        const int n_tiles = 2 * w.n_max;
#pragma unroll 1
        for (int n = -T::kSBufs; n < n_tiles; ++n) {
          const int t = n & 1, j = n >> 1;
          const int m = n + T::kSBufs, t2 = m & 1, j2 = m >> 1;
          const bool has_pv = n >= 0 && j < w.n(t);
          const bool has_s = m < n_tiles && j2 < w.n(t2);
          if (!has_pv && !has_s) continue;
          // ring tiles first used here (same order as the producer: V of the P V, then K of the S)
          bool waited = false;
          if (has_pv && kv_first_use(w, t, j)) {
            v_sh = ring++;
            wait_full(v_sh);
            waited = true;
          }
          if (has_s && kv_first_use(w, t2, j2)) {
            k_sh = ring++;
            wait_full(k_sh);
            waited = true;
          }
          if (waited) tc_fence_after();
          const int vbuf = v_sh % T::kNBuf, kbuf = k_sh % T::kNBuf;
          const uint32_t s_new = tmem_base + T::kTmemS + static_cast<uint32_t>((m % T::kSBufs) * kBlockN);
          const int gs = has_s ? (t2 == 0 ? gs_a++ : gs_b++) : 0;
          auto issue_next_s = [&]() {   // inside an elect block
            issue_s_at(s_new, set * 2 + (w.split ? 0 : t2), kbuf);
            tc_commit(bar_s_at(t2, gs));
            if (kv_last_use(w, t2, j2)) tc_commit(bar_empty + 8 * kbuf);
          };
          if (has_pv) {
            if (!ofree_done) {   // the first P V of an item overwrites O_t: the previous item's epilogues must have read it
              wait_ofree();
              ofree_done = true;
            }
            const int g = t == 0 ? gp_a++ : gp_b++;
            const uint32_t s_tile = tmem_base + T::kTmemS + static_cast<uint32_t>((n % T::kSBufs) * kBlockN);
            const uint32_t o_acc = tmem_base + T::kTmemO + static_cast<uint32_t>(t * kHeadDim);
            FA_TRACE_AT(2 + t, g, 0);
            if constexpr (T::kSplitP) {
              mbar_wait(bar_p_at(t, 0, g), par_at(g), TAG_P_FULL);
              tc_fence_after();
              FA_TRACE_AT(2 + t, g, 1);
              if (elect_one_sync()) issue_pv_at(o_acc, s_tile, vbuf, j > 0, 0, kKStepsSplit);
              __syncwarp();
              FA_TRACE_AT(2 + t, g, 2);
            }
            mbar_wait(bar_p_at(t, 1, g), par_at(g), TAG_P_FULL);
            tc_fence_after();
            FA_TRACE_AT(2 + t, g, 3);
            if (elect_one_sync()) {
              issue_pv_at(o_acc, s_tile, vbuf, j > 0, T::kSplitP ? kKStepsSplit : 0, kKStepsPV);
              tc_commit(bar_pv1 + 8 * t);     // O_t holds all of step j (the softmax rescales O_t only behind this)
              if (kv_last_use(w, t, j)) tc_commit(bar_empty + 8 * vbuf);
              if (j == w.n(t) - 1) tc_commit(bar_o + 8 * t);
              if (has_s) issue_next_s();
            }
            __syncwarp();
            FA_TRACE_AT(2 + t, g, 6);
          } else {
            if (elect_one_sync()) issue_next_s();
            __syncwarp();
          }
        }
      } else if (!w.split) {
        // ---- two Q tiles share every K/V tile: K_j then V_j in ring order ----
        const int r0 = ring;
        ring += 2 * w.n_max;
        wait_full(r0);
        tc_fence_after();
        if (seq == 0) FA_TRACE_MISC(2, 2);
        if (elect_one_sync()) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (w.n(t) > 0) {
              issue_s(t, set * 2 + t, r0 % T::kNBuf);
              tc_commit(bar_s + 8 * t);
            }
          }
          tc_commit(bar_empty + 8 * (r0 % T::kNBuf));
        }
        __syncwarp();
        if (seq == 0) FA_TRACE_MISC(2, 3);
        wait_ofree();
        for (int j = 0; j < w.n_max; ++j) {
          const int iv = r0 + 2 * j + 1, ik = r0 + 2 * j + 2;
          const int vbuf = iv % T::kNBuf, kbuf = ik % T::kNBuf;
          wait_full(iv);
#if FA_TRACE
          FA_TRACE_AT(2, steps_a, 5);      // V_j landed
#endif
          if (j + 1 < w.n_max) wait_full(ik);
#if FA_TRACE
          FA_TRACE_AT(3, steps_a, 5);      // K_(j+1) landed
#endif
          tc_fence_after();
#if FA_TRACE
          FA_TRACE_AT(2, steps_a, 7);      // K/V of this step landed (slot 0 of the same row = step start)
#endif
          // the ring slots of this step (V_j, K_(j+1)) are free once the last MMA that reads them has completed: slot B's
          // P V(j) and Q K^T(j+1) when slot B is the longer one (n1 == n_max: every causal and every full item) — released
          // from inside its step, in the elect block that issues those MMAs — else in a block of their own
          const bool rel_in_b = FA_OPT_RELEASE_IN_STEP != 0 && w.n1 == w.n_max;
          FA_STEP_UNROLL
          for (int t = 0; t < 2; ++t)
            if (j < w.n(t)) step(t, j, j == w.n(t) - 1, set * 2 + t, vbuf, kbuf, t == 1 && rel_in_b);
          if (!rel_in_b) {
            if (elect_one_sync()) {
              tc_commit(bar_empty + 8 * vbuf);
              if (j + 1 < w.n_max) tc_commit(bar_empty + 8 * kbuf);
            }
            __syncwarp();
          }
#if FA_TRACE
          FA_TRACE_AT(3, steps_b - 1, 7);  // ring slots of this step released
#endif
        }
      } else {
        // ---- split-KV: each slot has its own K/V tiles; ring order K_A0, K_B0, then V_A(j), K_A(j+1), V_B(j), K_B(j+1) ----
        const int qb = set * 2;
        const int ia = ring++;
        const int ib = w.n1 > 0 ? ring++ : ia;
        wait_full(ia);
        if (w.n1 > 0) wait_full(ib);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_s(0, qb, ia % T::kNBuf);
          tc_commit(bar_s);
          tc_commit(bar_empty + 8 * (ia % T::kNBuf));
          if (w.n1 > 0) {
            issue_s(1, qb, ib % T::kNBuf);
            tc_commit(bar_s + 8);
            tc_commit(bar_empty + 8 * (ib % T::kNBuf));
          }
        }
        __syncwarp();
        wait_ofree();
        for (int j = 0; j < w.n_max; ++j) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < w.n(t)) {
              const bool last = (j == w.n(t) - 1);
              const int iv = ring++;
              const int ik = last ? iv : ring++;
              wait_full(iv);
              if (!last) wait_full(ik);
              tc_fence_after();
              step(t, j, last, qb, iv % T::kNBuf, ik % T::kNBuf, true);
            }
          }
        }
      }
      // who will read TMEM in this item's epilogue: both slots of a 256-row item that had work; slot A alone (its own
      // accumulator and, for the merge, slot B's) in a 128-row item
      of_pend = w.single ? 1u : ((w.n0 > 0 ? 1u : 0u) | (w.n1 > 0 ? 2u : 0u));
    }
  } else if (kPrecise && warp >= 4) {
    // =========================== lo copies (precise instances: the warps of slot B) ===========================
    // lo = x - trunc_tf32(x) over the raw chunks of a tile, element by element: the lo chunks use the same swizzle as the raw
    // ones, so the copy is layout-agnostic.  Tiles are taken in the producer's order (first kNBuf K/V tiles, Q, the rest).
    const int ct = static_cast<int>(threadIdx.x) - 4 * 32;
    auto write_lo = [&](uint32_t tile, uint32_t bar) {
#pragma unroll 4
      for (int i = ct; i < T::kLoadBytes / 16; i += 128) {
        uint32_t x[4];
        ld_shared_v4(tile + i * 16, x[0], x[1], x[2], x[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          x[e] = __float_as_uint(__uint_as_float(x[e]) - __uint_as_float(x[e] & 0xFFFFE000u));
        st_shared_v4(tile + T::kLoOffset + i * 16, x[0], x[1], x[2], x[3]);
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to tcgen05.mma's operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    int ring = 0;
    uint32_t q_par = 0;
    for (int seq = 0;; ++seq) {
      const int item = next_item(seq);
      if (item < 0) break;
      const Item w = decode_item<kCausal>(p, item);
      if (w.n_max == 0) continue;
      const int qb = (seq % kQS) * 2;
      auto conv_q = [&]() {
        mbar_wait(bar_q + 8 * qb, q_par & 1u, TAG_Q_FULL);
        q_par ^= 1u;
        write_lo(sQ + T::q_tile(qb) * T::kTileBytes, bar_qconv + 8 * qb);
      };
      bool q_done = false;
      if (seq == 0) {
        conv_q();
        q_done = true;
      }
      for (int i = 0; i < 2 * w.n_max; ++i, ++ring) {
        if (!q_done && i == T::kNBuf) {
          conv_q();
          q_done = true;
        }
        const int buf = ring % T::kNBuf;
        mbar_wait(bar_full + 8 * buf, (ring / T::kNBuf) & 1, TAG_KV_FULL);
        write_lo(sKV + buf * T::kTileBytes, bar_conv + 8 * buf);
      }
      if (!q_done) conv_q();
    }
  } else {
    // =========================== softmax + epilogue (warps 0-7) ===========================
    const int t = warp >> 2;                       // tile slot of this warpgroup
    const int r = (warp & 3) * 32 + lane;          // row within the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS_fixed = tmem_base + lane_base + T::kTmemS + t * kBlockN;
    const uint32_t tO = tmem_base + lane_base + T::kTmemO + t * kHeadDim;
    const uint32_t tO_other = tmem_base + lane_base + T::kTmemO + (t ^ 1) * kHeadDim;
    const uint32_t tP1 = tmem_base + lane_base + T::kTmemP1 + t * T::kP1Cols;   // second piece of P (kEarlyS)
    const float c = p.scale_log2;
    const bool tracer = (warp & 3) == 0 && lane == 0;
    (void)tracer;
    int steps = 0;            // K/V steps of this slot so far, over all items (parity of bar_s / bar_p)
    uint32_t o_par = 0;       // bit u: parity of the next bar_o[u] phase (one phase per item in which slot u has work)
    // Instances with two Q sets: the thread that issued an item's O store does not wait for the store to have read the
    // staging buffer right away (~800 cycles in which its warp, and with it the slot's next softmax step, would stall:
    // timeline of C3, profiles/r01_trace_small_shapes.txt) — the buffer is not needed again before the item after next.
    // It waits, and releases the buffer to the producer, after the first softmax step of the next item instead.
    constexpr bool kLateQFree = FA_OPT_LATE_QFREE != 0 && kQS > 1;
    const bool store_thread = (warp & 3) == 0 && lane == 0;
    int pend_qfree = -1;      // store_thread only: Q buffer whose O store has been issued but not yet waited for
    auto flush_qfree = [&]() {
      if (kLateQFree && store_thread && pend_qfree >= 0) {
        tma_store_wait_read();
        mbar_arrive(bar_qfree + 8 * pend_qfree);
        pend_qfree = -1;
      }
    };

    // exp2 of keys [i0, i0 + 32) of the S row held in s[], in place, with the partial row sums in l0..l3
    auto exp_chunk = [&](float* s, const int i0, const float neg_mc, float& l0, float& l1, float& l2, float& l3) {
#pragma unroll
      for (int i = i0; i < i0 + 32; i += 4) {
        // which of these two element pairs take the polynomial route
        const bool kPoly01 = T::kPacked && ((i >> 1) % T::kPolyDen) < T::kPolyNum;
        const bool kPoly23 = T::kPacked && (((i >> 1) + 1) % T::kPolyDen) < T::kPolyNum;
        bool packed = false;
#if FA_OPT_F2
        if constexpr (T::kPacked) {
          packed = true;
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
          const float2 s01 = fadd2(make_float2(l0, l1), make_float2(s[i], s[i + 1]));
          const float2 s23 = fadd2(make_float2(l2, l3), make_float2(s[i + 2], s[i + 3]));
          l0 = s01.x; l1 = s01.y; l2 = s23.x; l3 = s23.y;
        }
#endif
        if (!packed) {
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
          if constexpr (kTF32 && !T::kComp) {
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
      }
    };

    for (int seq = 0;; ++seq) {
      const int item = next_item(seq);
      if (item < 0) break;
      const Item w = decode_item<kCausal>(p, item);
      const int set = seq % kQS;
      const int n_mine = w.n(t);
      const int q_row = (t == 0 ? w.row0 : w.row1) + r;
      const int kv_first = w.kv_base + (t == 0 ? 0 : w.kv_first1);

      float m = -INFINITY;  // running (possibly stale) row max, in raw q.k units
      float l = 0.f;        // running row sum of exp2((s - m) * c)

      for (int j = 0; j < n_mine; ++j) {
        const int g = steps++;
        if (tracer) FA_TRACE_AT(t, g, 0);
        mbar_wait(bar_s_at(t, g), par_at(g), TAG_S_FULL);
        tc_fence_after();
        // this step's S tile (P overwrites it in place): fixed per slot, or tile number % 3 of the rotating buffers
        const uint32_t tS = T::kRotS ? tmem_base + lane_base + T::kTmemS + static_cast<uint32_t>(((2 * j + t) % T::kSBufs) * kBlockN) : tS_fixed;
        if (tracer) FA_TRACE_AT(t, g, 1);
        float s[128];
        // masking: key kv0 + i is visible iff i <= limit
        const int kv0 = (kv_first + j) * kBlockN;
        int limit = p.n_k - 1 - kv0;
        if (kCausal) limit = min(limit, q_row + p.causal_offset - kv0);
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
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
        const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
        if (tracer) FA_TRACE_AT(t, g, 2);

        if (j == 0) {
          m = m_new;
        } else {
          // lazy rescale: only move the reference max when it grew by more than 2^kRescaleThreshold
          const bool need = (m_new - m) * c > kRescaleThreshold;   // (-inf -> finite) gives +inf -> true
          if (__any_sync(0xffffffffu, need)) {
            if constexpr (T::kEarlyS || T::kRotS) {
              // S_t(j) no longer implies that all of P_t(j-1) V has retired: wait for its second piece before touching O_t
              mbar_wait(bar_pv1 + 8 * t, (g - 1) & 1, TAG_PV1);
              tc_fence_after();
            }
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
        const float neg_mc = T::kComp ? fmaf(-m_safe, c, kTf32CompLog2) : -m_safe * c;
        // P = exp2(s*c - m*c) in two pieces (keys [0, kSplitKeys) and the rest); each piece is written to TMEM and handed
        // to the MMA warp as soon as it is complete, so most of P*V runs under the exps of the last piece and only a
        // short P*V remains between the last arrival and the next Q*K^T.
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        constexpr int kChunks0 = T::kSplitKeys / 32;   // 32-key chunks in the first piece
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int cc = (h == 0 ? 0 : kChunks0); cc < (h == 0 ? kChunks0 : 4); ++cc) {
            exp_chunk(s, cc * 32, neg_mc, l0, l1, l2, l3);
            const bool own = T::kEarlyS && h == 1;   // second piece in its own columns
            if (own && cc == kChunks0 && g > 0) {
              // the previous step's second-piece P*V must have read these columns (long done: it was issued a whole
              // softmax step ago)
              mbar_wait(bar_pv1 + 8 * t, (g - 1) & 1, TAG_PV1);
              tc_fence_after();
            }
            if constexpr (kTF32) {
              tmem_st32(own ? tP1 + (cc - kChunks0) * 32 : tS + cc * 32, reinterpret_cast<uint32_t*>(&s[cc * 32]));
              if constexpr (kPrecise) {   // what kind::tf32 drops from P goes to the lo columns
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  uint32_t lo[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float x = s[cc * 32 + h2 * 16 + i];
                    lo[i] = __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
                  }
                  tmem_st16(tS + T::kTmemPLo + cc * 32 + h2 * 16, lo);
                }
              }
            } else {
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_16x2<kF16>(s[cc * 32 + 2 * i], s[cc * 32 + 2 * i + 1]);
              tmem_st16(own ? tP1 + (cc - kChunks0) * 16 : tS + cc * 16, pk);
            }
          }
          if (T::kSplitP || h == 1) {   // one-piece instances arrive once, on the second barrier
            if (tracer) FA_TRACE_AT(t, g, 3 + 2 * h);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p_at(t, h, g));
            if (tracer) FA_TRACE_AT(t, g, 4 + 2 * h);
          }
        }
        l += (l0 + l1) + (l2 + l3);
        if (j == 0) flush_qfree();
      }
      flush_qfree();   // (a slot without K/V steps in this item)
      if constexpr (T::kComp) l *= kTf32CompInv;   // the sums were taken over P*(1+eps)

      // ---- epilogue: O/l -> swizzled SMEM (reusing this slot's Q buffer) -> TMA store; LSE -> global ----
      if (tracer && seq == 0) FA_TRACE_MISC(t, 2);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        // own accumulator final; in split mode slot A also needs the partner's (it merges O_B).  A completion that is
        // not waited for here still happens: keep its parity count in step.
        if (w.n(u) > 0) {
          if (u == t || (w.split && t == 0)) mbar_wait(bar_o + 8 * u, (o_par >> u) & 1u, TAG_O_FINAL);
          o_par ^= 1u << u;
        }
      }
      tc_fence_after();
      if (tracer && seq == 0) FA_TRACE_MISC(t, 3);
      // scale of this slot's accumulator and of the partner's partial (split-KV tail items only) in the final O
      // A row that has seen no visible key (m still -inf: causal with n_q > n_k masks whole rows inside partly visible
      // tiles) is O = 0, LSE = -inf like the oracle's; its l may hold 2^-126 crumbs from the polynomial exp2 route, which
      // clamps its argument, so the test is on m, not on l alone.
      const bool seen = n_mine > 0 && m != -INFINITY && l > 0.f;
      float f_self = seen ? 1.0f / l : 0.f;
      float f_other = 0.f;
      float lse_val = seen ? m * p.scale + logf(l) : -INFINITY;
      // Who writes output rows.  A slot that got a Q tile (=> it has keys, and rows below n_q) stages its O tile in that
      // buffer, stores it with TMA and then releases the buffer to the producer.  A slot without one has no staging buffer
      // either — its Q buffer may already be receiving the next item's tile — so rows that exist but see no key at all
      // (causal, n_q > n_k) are written as zeros with plain global stores, and slots without rows (slot B of a 128-row item;
      // rows past n_q: slot B of the last 256-row block when n_q % 256 is in [1, 128], the dead second item of a one-slot
      // instance) write nothing.
      const bool stores = w.loads_q(t);
      const bool zero_rows = !stores && !(w.single && t == 1) && q_row < p.n_q;
      const bool merge = w.split && w.n1 > 0;
      if (merge) {
        if (t == 1) {
          st_shared_b32(s_ml + r * 4, __float_as_uint(m));
          st_shared_b32(s_ml + (kBlockM + r) * 4, __float_as_uint(l));
          named_bar_arrive(kBarMerge, 256);
        } else {
          named_bar_sync(kBarMerge, 256);
          const float m_b = __uint_as_float(ld_shared_b32(s_ml + r * 4));
          const float l_b = __uint_as_float(ld_shared_b32(s_ml + (kBlockM + r) * 4));
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
      // accumulate mode: weights of the earlier partial and of this launch's result in the merged row
      float w_acc = 0.f;
      const float* acc_row = nullptr;
      if (p.acc_o != nullptr && stores) {
        float w_new = 1.f;
        if (q_row < p.n_q) {
          const int64_t row_idx = (static_cast<int64_t>(w.batch) * p.heads + w.head) * p.n_q + q_row;   // (never with kv_splits)
          acc_row = p.acc_o + row_idx * p.head_dim;
          const float la = p.acc_lse[row_idx], lb = lse_val;
          const float mx = fmaxf(la, lb);
          if (mx == -INFINITY) {
            w_new = 0.f;
          } else {
            const float ea = expf(la - mx), eb = expf(lb - mx);
            const float inv = 1.0f / (ea + eb);
            w_acc = ea * inv;
            w_new = eb * inv;
            lse_val = mx + logf(ea + eb);
          }
        }
        f_self *= w_new;
        f_other *= w_new;
      }
      const uint32_t stage = sQ + T::q_tile(set * 2 + t) * T::kTileBytes;   // kDChunks boxes of 16 KB (slot B of a one-slot
                                                                            // instance never stores: `stores` below)
      const uint32_t row_off = r * 128;
      const uint32_t sw = r & 7;
      constexpr int kRounds = T::kOChunks / T::kDChunks;   // 1, or 2 for bf16-in / fp32-out
      constexpr int kColsPerChunk = T::kOutElemsPerChunk;  // 32 (fp32) or 64 (bf16)
      constexpr int kLastCol0 = (kRounds * T::kDChunks - 1) * kColsPerChunk + (kColsPerChunk / 32 - 1) * 32;
      if (!stores) {
        if (zero_rows) {
          uint8_t* row = static_cast<uint8_t*>(p.o_ptr) +
                         (static_cast<int64_t>(w.obatch) * p.o_sb + static_cast<int64_t>(w.head) * p.o_sh + static_cast<int64_t>(q_row) * p.o_sn) * T::kOutSize;
          for (int b = 0; b < p.o_row_bytes; b += 16) *reinterpret_cast<uint4*>(row + b) = make_uint4(0u, 0u, 0u, 0u);
          if (p.lse != nullptr) p.lse[(static_cast<int64_t>(w.obatch) * p.heads + w.head) * p.n_q + q_row] = -INFINITY;
        }
      } else {
#pragma unroll
        for (int round = 0; round < kRounds; ++round) {
#pragma unroll
          for (int ch = 0; ch < T::kDChunks; ++ch) {
            const int col0 = (round * T::kDChunks + ch) * kColsPerChunk;
#pragma unroll
            for (int half = 0; half < kColsPerChunk / 32; ++half) {
              const int cbase = col0 + half * 32;
              uint32_t o[32];
              if (n_mine > 0) {
                tmem_ld32(tO + cbase, o);
                tc_wait_ld();
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = 0u;
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f_self);
              if (merge) {
                uint32_t ob[32];
                tmem_ld32(tO_other + cbase, ob);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(fmaf(__uint_as_float(ob[i]), f_other, __uint_as_float(o[i])));
              }
              if (acc_row != nullptr) {
                // this row's 32 columns of the earlier partial, straight from global memory (L2): 512 B of one row per
                // thread, once per item — noise next to the K/V stream of the item's main loop
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  if (cbase + 4 * g < p.head_dim) {
                    const float4 a = *reinterpret_cast<const float4*>(acc_row + cbase + 4 * g);
                    o[4 * g] = __float_as_uint(fmaf(a.x, w_acc, __uint_as_float(o[4 * g])));
                    o[4 * g + 1] = __float_as_uint(fmaf(a.y, w_acc, __uint_as_float(o[4 * g + 1])));
                    o[4 * g + 2] = __float_as_uint(fmaf(a.z, w_acc, __uint_as_float(o[4 * g + 2])));
                    o[4 * g + 3] = __float_as_uint(fmaf(a.w, w_acc, __uint_as_float(o[4 * g + 3])));
                  }
                }
              }
              if (cbase == kLastCol0 && n_mine > 0) {
                // last TMEM read of this item: the MMA warp may overwrite O_t (and O_B after a merge) with the next item's
                // first P*V.  Slots without work read nothing and do not arrive (see wait_ofree in the MMA warp).
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ofree + 8 * t);
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
                               pack_16x2<kF16>(__uint_as_float(o[8 * g]), __uint_as_float(o[8 * g + 1])),
                               pack_16x2<kF16>(__uint_as_float(o[8 * g + 2]), __uint_as_float(o[8 * g + 3])),
                               pack_16x2<kF16>(__uint_as_float(o[8 * g + 4]), __uint_as_float(o[8 * g + 5])),
                               pack_16x2<kF16>(__uint_as_float(o[8 * g + 6]), __uint_as_float(o[8 * g + 7])));
                }
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + t, 128);
          if (tracer && seq == 0) FA_TRACE_MISC(t, 4);
          if ((warp & 3) == 0 && lane == 0) {
#pragma unroll
            for (int ch = 0; ch < T::kDChunks; ++ch)
              tma_store_4d(&tm_o, stage + ch * kChunkBytes, (round * T::kDChunks + ch) * kColsPerChunk, w.row0 + t * kBlockM,
                           w.head, w.obatch);
            tma_store_commit();
            if (kLateQFree && round + 1 == kRounds) {
              pend_qfree = set * 2 + t;
            } else {
              tma_store_wait_read();
              if (round + 1 == kRounds) mbar_arrive(bar_qfree + 8 * (set * 2 + t));   // staging buffer read out: Q may land here again
            }
            if (seq == 0) FA_TRACE_MISC(t, 5);
          }
          if (round + 1 < kRounds) named_bar_sync(1 + t, 128);
        }
        // after the staging stores: a global store ahead of fence.proxy.async would make that fence wait for it
        if (p.lse != nullptr && q_row < p.n_q)
          p.lse[(static_cast<int64_t>(w.obatch) * p.heads + w.head) * p.n_q + q_row] = lse_val;
      }
    }
    flush_qfree();
    if ((warp & 3) == 0 && lane == 0) {
      tma_store_wait_all();
      FA_TRACE_MISC(t, 6);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
  if (p.work_counter != nullptr && threadIdx.x == 0) {
    // the last CTA to finish re-arms the counters for the next launch that uses them
    __threadfence();
    const unsigned int done = atomicAdd(p.work_counter + 1, 1u);
    if (done == gridDim.x - 1) {
      p.work_counter[0] = 0u;
      p.work_counter[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace fa
