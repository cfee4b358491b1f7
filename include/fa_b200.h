/*
 * fa_b200.h — C-ABI of the B200-native FlashAttention forward (libfa_b200.so).
 *
 * Plain pointers and sizes only: no torch types, no C++ in the signatures.  Every entry
 * point states the reference interface (file:line into kilianhae/FlashAttention.C) it
 * stands in for.  All device entry points are asynchronous on `stream` (a cudaStream_t
 * passed as void*; NULL = the legacy default stream) and return 0 on success or a negative
 * fa_status code; fa_strerror() turns the code into text.  Nothing here ever falls back to
 * a CPU path: without a usable sm_100 device the calls return FA_ERR_NO_DEVICE.
 *
 * Shapes: Q is [batch, heads, n_q, head_dim], K and V are [batch, heads, n_k, head_dim],
 * O is [batch, heads, n_q, head_dim], LSE (optional) is [batch, heads, n_q] fp32 holding
 * log(sum_j exp(scale * q.k_j)) per row.  The reference's 3-D [B*H, N, d] tensors
 * (src/flashattention.cu:603-606) are the case batch = 1, heads = B*H.
 */
#ifndef FA_B200_H
#define FA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_B200_VERSION 102 /* major*100 + minor */

/* element type of Q, K, V (and of O unless stated otherwise) */
enum fa_dtype {
  FA_F32 = 0,  /* fp32 in HBM, contractions run as tcgen05 kind::tf32 (fp32 accumulate) */
  FA_BF16 = 1, /* bf16 in HBM, contractions run as tcgen05 kind::f16  (fp32 accumulate) */
  FA_F16 = 2   /* IEEE fp16 in HBM, same kind::f16 instruction with fp16 operands (fp32 accumulate) */
};

enum fa_status {
  FA_OK = 0,
  FA_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, unsupported head_dim ... */
  FA_ERR_NO_DEVICE = -2,     /* no CUDA device, or the device is not sm_100 */
  FA_ERR_CUDA = -3,          /* a CUDA runtime/driver call failed; see fa_last_cuda_error() */
  FA_ERR_UNSUPPORTED = -4,   /* combination not implemented (e.g. misaligned strides) */
  FA_ERR_ALIGNMENT = -5      /* pointer or stride not 16-byte aligned (TMA requirement) */
};

/* which kernel family served the most recent launch on this thread (fa_last_impl) */
enum fa_impl {
  FA_IMPL_NONE = 0,
  FA_IMPL_TCGEN05 = 1,  /* TMA + tcgen05.mma + TMEM warp-specialised kernel (the product) */
  FA_IMPL_SIMT = 2      /* CUDA-core general-shape kernel (odd head dims; the on-GPU checker) */
};

enum fa_flags {
  /* Every 128-row query tile is computed by the same sequence of operations whatever else is in the launch, so a
   * (batch, head) slice gives bit-identical results alone, inside a larger batch, or on another rank of a B x H
   * sharded job.  Off by default: the scheduler then splits the K/V range of the last partial wave's tiles over two
   * tile slots (split-KV with a log-sum-exp merge), which is faster for small grids but rounds differently. */
  FA_FLAG_BATCH_INVARIANT = 1,
  /* fp32 inputs only: fp32-grade contractions instead of plain tf32 ("3xTF32").  Every operand x is fed to the tensor core
   * twice — as it is (kind::tf32 reads hi = the top 19 bits) and as lo = x - hi — and both contractions run hi*hi + lo*hi +
   * hi*lo into one fp32 accumulator; exp2 is MUFU only (no polynomial).  Head dims <= 64 run on the tcgen05 kernel (about
   * 3x the tensor work of the default), larger ones on the fp32 CUDA-core kernel.  This is what lets the llm.c harness
   * keep its validate_result(out, 1e-4f) gate (src/llm.c/attention_forward.cu:1262), which plain tf32 (10-bit mantissa:
   * an output row that copies one V row is already off by up to 2.4e-4) cannot meet; the attention_forward[6] shims set
   * it.  FA_B200_PRECISE=1 in the environment sets it for every fp32 call of the process. */
  FA_FLAG_PRECISE = 2
};

/* strided problem description.  Strides are in ELEMENTS; the head_dim axis is contiguous. */
typedef struct fa_params {
  const void* q; const void* k; const void* v;
  void* o;                 /* same dtype as q unless o_f32 != 0 */
  float* lse;              /* optional [batch, heads, n_q] contiguous fp32; may be NULL */
  int64_t batch, heads, n_q, n_k;
  int32_t head_dim;        /* tcgen05 path: fp32 d <= 128 (d % 4 == 0), bf16 / fp16 d <= 256 (d % 8 == 0) — kernel instances exist
                              for 128-, 256- and 512-byte rows, head dims in between are zero-padded by TMA; fp32 d up to
                              256 (d % 8 == 0) runs on the CUDA-core kernel */
  int32_t dtype;           /* enum fa_dtype */
  int32_t causal;          /* 0 / 1.  Causal is bottom-right aligned: key j visible to row i iff j <= i + (n_k - n_q) */
  int32_t o_f32;           /* bf16 / fp16 inputs only: write O as fp32 (used by the ring merge) */
  float scale;             /* multiplies q.k before the softmax; the reference's torch path uses 1.0f
                              (src/flashattention.cu:593), its llm.c path 1/sqrt(d) (src/llm.c/attention_forward.cu:1123) */
  int64_t q_stride_b, q_stride_h, q_stride_n;
  int64_t k_stride_b, k_stride_h, k_stride_n;
  int64_t v_stride_b, v_stride_h, v_stride_n;
  int64_t o_stride_b, o_stride_h, o_stride_n;
  int32_t impl;            /* 0 = automatic; FA_IMPL_TCGEN05 / FA_IMPL_SIMT force a kernel family (tests) */
  int32_t flags;           /* bit set of enum fa_flags (0 = defaults) */
  /* Accumulate mode (tcgen05 kernel only; both NULL = plain forward).  o_acc: fp32 [batch, heads, n_q, head_dim]
   * contiguous, lse_acc: fp32 [batch, heads, n_q] — an earlier, normalised partial result of the same queries over OTHER
   * keys.  The kernel folds it into its own result by the log-sum-exp rule inside its epilogue (what fa_merge_partials
   * does as a separate pass): O and LSE come out merged.  o may be o_acc and lse may be lse_acc (in place).  This is a
   * ring-forward step: the merge costs no launch and no extra pass over O, and the last step can write 16-bit O directly. */
  const float* o_acc;
  const float* lse_acc;
  /* Grouped-query / multi-query attention: K and V have kv_heads heads ([batch, kv_heads, n_k, head_dim], their own strides) and
   * query head h reads K/V head h / (heads / kv_heads).  0 (or == heads) = one K/V head per query head.  heads % kv_heads == 0.
   * The K/V tiles of a group are fetched once per query head but by neighbouring items, so they are shared in L2. */
  int64_t kv_heads;
} fa_params;

/*
 * fa_forward — O = softmax(scale * Q K^T [+ causal mask]) V on contiguous [batch, heads, n, d].
 * Replaces the launcher + kernel pair run_flash_tiled_coarse[_causal] -> flash_tiled_coarse[_causal]
 * (src/flashattention.cu:590-602, 139-355, 359-579) behind forward() (src/flashattention.cu:603-617).
 * Unlike the reference it does not synchronise the device and it reports launch errors.
 */
int fa_forward(const void* q, const void* k, const void* v, void* o, float* lse,
               int64_t batch, int64_t heads, int64_t n_q, int64_t n_k, int32_t head_dim,
               float scale, int32_t causal, int32_t dtype, void* stream);

/* fa_forward_ex — the same through an explicit strided description (packed / head-interleaved layouts). */
int fa_forward_ex(const fa_params* p, void* stream);

/*
 * fa_forward_packed_qkv — llm.c layout: inp is (B, T, 3, NH, hs) fp32, out is (B, T, NH, hs) fp32.
 * Replaces permute_kernel -> flashattention -> unpermute_kernel -> D2D copy inside attention_forward6
 * (src/llm.c/attention_forward.cu:1106-1179, 519-565, 881-1104): the strided views go straight into
 * the TMA descriptors, so there is no temporary, no cudaMalloc and no layout pass.
 */
int fa_forward_packed_qkv(const float* inp, float* out, float* lse,
                          int32_t B, int32_t T, int32_t NH, int32_t hs,
                          float scale, int32_t causal, void* stream);
/* the same with a bit set of enum fa_flags (FA_FLAG_PRECISE is what the attention_forward[6] shims pass) */
int fa_forward_packed_qkv_ex(const float* inp, float* out, float* lse,
                             int32_t B, int32_t T, int32_t NH, int32_t hs,
                             float scale, int32_t causal, int32_t flags, void* stream);

/*
 * fa_backward — gradients of O = softmax(scale Q K^T [+ causal]) V: dQ, dK, dV from (Q, K, V, O, LSE, dO).  The reference is
 * forward only (README.md:33 lists what it leaves open); this is the step after its operator for a training caller, on the same
 * tcgen05 / TMEM / TMA machinery: a statistics pass (D = rowsum(dO * O)), one launch that owns 128 keys per CTA and accumulates
 * dK, dV over the query tiles, and one that owns 128 query rows per CTA and accumulates dQ over the key tiles — no atomics, so
 * results are bit-reproducible.  bf16 and fp16, head_dim <= 128 (% 8 == 0); FA_F32 returns FA_ERR_UNSUPPORTED.  LSE is the
 * forward's ([batch, heads, n_q] contiguous fp32); O and dO are [batch, heads, n_q, head_dim]; dQ is shaped like Q, dK and dV
 * like K and V ([batch, kv_heads, n_k, head_dim]: with grouped K/V heads they are summed over the group's query heads).  Strides
 * are in ELEMENTS with the head_dim axis contiguous; gradient pointers and strides must be 16-byte aligned.
 */
typedef struct fa_bwd_params {
  const void* q; const void* k; const void* v; const void* o; const void* d_o;
  const float* lse;
  void* dq; void* dk; void* dv;
  int64_t batch, heads, kv_heads /* 0 = heads */, n_q, n_k;
  int32_t head_dim, dtype, causal;
  float scale;
  int64_t q_stride_b, q_stride_h, q_stride_n;
  int64_t k_stride_b, k_stride_h, k_stride_n;
  int64_t v_stride_b, v_stride_h, v_stride_n;
  int64_t o_stride_b, o_stride_h, o_stride_n;
  int64_t do_stride_b, do_stride_h, do_stride_n;
  int64_t dq_stride_b, dq_stride_h, dq_stride_n;
  int64_t dk_stride_b, dk_stride_h, dk_stride_n;
  int64_t dv_stride_b, dv_stride_h, dv_stride_n;
} fa_bwd_params;
int fa_backward(const fa_bwd_params* p, void* stream);

/*
 * fa_forward_host — the same operator with HOST buffers: copies Q, K, V to the device, runs
 * fa_forward, copies O back and synchronises.  This is the end-to-end call bench.py times
 * ("e2e"); it mirrors what bench_flashattention.py:31-33,70 does around forward().
 * Device scratch is cached inside the library and reused between calls.
 */
int fa_forward_host(const void* q_host, const void* k_host, const void* v_host, void* o_host,
                    int64_t batch, int64_t heads, int64_t n_q, int64_t n_k, int32_t head_dim,
                    float scale, int32_t causal, int32_t dtype);

/*
 * fa_merge_partials — log-sum-exp merge of two attention partials over disjoint key sets, in place:
 *   lse = log(exp(lse_acc) + exp(lse_new));  o_acc = o_acc*exp(lse_acc-lse) + o_new*exp(lse_new-lse)
 * o_acc / o_new are fp32 [rows, head_dim] contiguous, lse_* fp32 [rows].  Used by the ring
 * (sequence-partitioned) forward.  The reference plumbs `out_l` (src/flashattention.cu:140, 609)
 * but never writes it; this is the consumer it was meant for.
 */
int fa_merge_partials(float* o_acc, float* lse_acc, const float* o_new, const float* lse_new,
                      int64_t rows, int32_t head_dim, void* stream);

/* fa_cast_f32 — final cast of the ring accumulator: [n] fp32 -> dtype (FA_BF16 or FA_F16).  fa_cast_f32_to_bf16 is the
 * bf16 case under its original name. */
int fa_cast_f32(const float* src, void* dst, int64_t n, int32_t dtype, void* stream);
int fa_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/*
 * Peer-to-peer staging for the sequence-partitioned (ring) forward, one process per GPU (not in the reference, which is
 * single-GPU).  fa_p2p_alloc makes a device buffer and its 64-byte CUDA IPC handle; the other ranks fa_p2p_open the handle
 * (exchanged by the host plumbing, e.g. torch.distributed all_gather) and pull the K/V shard they need next with
 * fa_copy_async: a device-to-device copy that the copy engines run over NVLink without occupying an SM, so it overlaps the
 * persistent attention kernel of the current step.  The caller orders the accesses (a barrier after the owners filled
 * their buffers and one before they are reused).
 */
int fa_p2p_alloc(int64_t bytes, void** ptr, uint8_t handle[64]);
int fa_p2p_open(const uint8_t handle[64], void** peer_ptr);   /* in a process other than the exporter's */
int fa_p2p_close(void* peer_ptr);
int fa_p2p_free(void* ptr);
int fa_copy_async(void* dst, const void* src, int64_t bytes, void* stream);   /* src or dst may be a peer mapping */

/*
 * Reference-named shims (same argument order and meaning as the reference symbols).
 *
 * run_flash_tiled_coarse / _causal: test.cu:591-603 (torch-less launchers; note the O, K, Q, V order),
 *   [batch, seq, 64] fp32, scale 1.0, synchronous like the reference.
 * attention_forward6 / attention_forward: src/llm.c/attention_forward.cu:1106-1109 and 1183-1211.
 *   Only kernel_num 6 (the author's flash kernel) is served; the llm.c comparison kernels 1-5 are out of
 *   scope and, like an invalid number in the reference (1207-1209), print a message and exit(1).
 *   Errors print and exit(EXIT_FAILURE) like cudaCheck (src/llm.c/common.h:16-23).
 */
void run_flash_tiled_coarse(float* O, float* K_d, float* Q_d, float* V_d, int batch_size, int seq_len);
void run_flash_tiled_coarse_causal(float* O, float* K_d, float* Q_d, float* V_d, int batch_size, int seq_len);
void attention_forward6(float* out, const float* inp, int B, int T, int C, int NH, const int block_size);
void attention_forward(int kernel_num, float* out, float* vaccum, float* qkvr, float* preatt, float* att,
                       const float* inp, int B, int T, int C, int NH, const int block_size);

/*
 * fa_query_instance — which kernel serves (dtype, head_dim); needs no device.  > 0: the tcgen05 kernel instance (its head dim:
 * 32 / 64 / 128 for fp32, 64 / 128 / 256 for bf16 and fp16; a smaller head_dim is zero-padded onto it by TMA); 0: the CUDA-core
 * kernel (fp32 head dims in (128, 256], multiples of 8); < 0: FA_ERR_UNSUPPORTED / FA_ERR_INVALID_ARG.  The reference fixes the
 * head dim at compile time (`# define d 64`, src/flashattention.cu:15) and supports nothing else.
 */
int fa_query_instance(int32_t dtype, int32_t head_dim);

/* diagnostics */
const char* fa_strerror(int status);
const char* fa_last_cuda_error(void); /* text of the last CUDA failure seen by this thread ("" if none) */
int fa_last_impl(void);               /* enum fa_impl of the last successful launch on this thread */
int fa_version(void);
int64_t fa_launch_count(void);        /* number of kernels this library has launched in this process */
/* Every mbarrier wait in the kernels is bounded; on expiry the CTA records where it was stuck and traps, so a pipeline
 * bug surfaces as a CUDA error instead of a hung GPU.  out = {barrier tag (0 = never expired), blockIdx.x, threadIdx.x,
 * phase parity} of the first expiry in this process (kept in host-mapped memory, readable after the failed launch). */
int fa_watchdog_info(uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H */
