// fa_api.cu — the C-ABI of libfa_b200.so (see include/fa_b200.h): argument validation, TMA tensor-map
// construction, kernel-instance dispatch, the host-buffer (e2e) entry and the reference-named shims.
#include "../../include/fa_b200.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "fa_fwd_sm100.cuh"
#include "fa_simt.cuh"
#include "fa_bwd_sm100.cuh"

namespace {

thread_local char t_cuda_err[512] = "";
thread_local int t_last_impl = FA_IMPL_NONE;
std::atomic<int64_t> g_launches{0};

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(t_cuda_err, sizeof(t_cuda_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return FA_ERR_CUDA;
}
#define FA_CUDA(call)                                    \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// ---- device probe (cached per device) ----
struct DevInfo { int checked = 0; int major = 0; int minor = 0; int sms = 0; };
DevInfo g_dev[64];
std::mutex g_dev_mu;

int probe_device(int* major_out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return FA_ERR_NO_DEVICE;
  }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& di = g_dev[dev];
  if (!di.checked) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      cudaGetLastError();
      return FA_ERR_NO_DEVICE;
    }
    di.major = prop.major; di.minor = prop.minor; di.sms = prop.multiProcessorCount; di.checked = 1;
  }
  *major_out = di.major;
  return FA_OK;
}

int current_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  return g_dev[dev].checked ? g_dev[dev].sms : 148;
}

// ---- work counters of the persistent kernel: {next item, CTAs finished}; the last CTA of a launch zeroes its pair, so a
// pair is reusable as soon as the launch that used it has finished.  A pair belongs to ONE (device, stream): launches on a
// stream run one after the other, so they can share it; launches on different streams may overlap and never do.  A launch
// that is being captured into a CUDA graph gets no pair at all (static item stride): a graph keeps the pointer for good and
// may be replayed on any stream, concurrently with eager launches or with another instantiation of itself.
constexpr int kCounterPool = 1024;   // streams per device that get dynamic scheduling; beyond that: static stride
struct CounterPool {
  unsigned int* base = nullptr;
  int used = 0;
  std::unordered_map<cudaStream_t, int> slot_of;
};
CounterPool g_counters[64];
std::mutex g_counters_mu;

// host-mapped record of the first mbarrier watchdog expiry (see ptx.cuh); one per process, shared by all devices
unsigned int* g_wd_host = nullptr;
bool g_wd_installed[64] = {};
std::mutex g_wd_mu;
void install_watchdog_record(int dev) {
  if (dev < 0 || dev >= 64 || g_wd_installed[dev]) return;
  std::lock_guard<std::mutex> lk(g_wd_mu);
  if (g_wd_installed[dev]) return;
  if (!g_wd_host) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_wd_host), 4 * sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      g_wd_host = nullptr;
    } else {
      memset(g_wd_host, 0, 4 * sizeof(unsigned int));
    }
  }
  if (g_wd_host) {
    unsigned int* dptr = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), g_wd_host, 0) == cudaSuccess)
      cudaMemcpyToSymbol(fa::g_fa_watchdog_host, &dptr, sizeof(dptr));
    cudaGetLastError();
  }
  g_wd_installed[dev] = true;
}

int next_work_counter(cudaStream_t st, unsigned int** out) {
  *out = nullptr;
  int dev = 0;
  FA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return FA_ERR_NO_DEVICE;
  install_watchdog_record(dev);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  FA_CUDA(cudaStreamIsCapturing(st, &cap));
  if (cap != cudaStreamCaptureStatusNone) return FA_OK;   // captured launch: static stride
  std::lock_guard<std::mutex> lk(g_counters_mu);
  CounterPool& cp = g_counters[dev];
  if (!cp.base) {
    unsigned int* b = nullptr;
    FA_CUDA(cudaMalloc(&b, kCounterPool * 2 * sizeof(unsigned int)));
    FA_CUDA(cudaMemset(b, 0, kCounterPool * 2 * sizeof(unsigned int)));   // synchronous w.r.t. the host: done before any launch
    cp.base = b;
  }
  auto it = cp.slot_of.find(st);
  if (it == cp.slot_of.end()) {
    if (cp.used == kCounterPool) return FA_OK;   // pool exhausted: this stream runs with the static stride
    it = cp.slot_of.emplace(st, cp.used++).first;
  }
  *out = cp.base + 2 * it->second;
  return FA_OK;
}

// ---- workspace of split-KV launches (partial O and LSE of every run).  Like the work counters it belongs to one (device, stream):
// launches on a stream run one after the other and reuse it (grow-only), launches on different streams have their own.  A launch
// that is being captured gets a stream-ordered allocation of its own instead (cudaMallocAsync / cudaFreeAsync become graph nodes):
// a graph may be replayed on any stream.  (Outside capture cudaMallocAsync per call costs ~150 us with the default pool settings.)
struct SplitWs { void* ptr = nullptr; size_t cap = 0; };
std::unordered_map<cudaStream_t, SplitWs> g_split_ws[64];
std::vector<void*> g_split_ws_retired;
std::mutex g_split_ws_mu;

int split_workspace(cudaStream_t st, size_t bytes, void** out, void** async_owned) {
  *async_owned = nullptr;
  int dev = 0;
  FA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return FA_ERR_NO_DEVICE;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  FA_CUDA(cudaStreamIsCapturing(st, &cap));
  if (cap != cudaStreamCaptureStatusNone) {
    FA_CUDA(cudaMallocAsync(out, bytes, st));
    *async_owned = *out;
    return FA_OK;
  }
  std::lock_guard<std::mutex> lk(g_split_ws_mu);
  SplitWs& w = g_split_ws[dev][st];
  if (w.cap < bytes) {
    // grow by 1.5x; the outgrown buffer is kept (a launch enqueued earlier on this stream may still be using it, and a
    // synchronise-and-free here would stall the caller): growth is geometric, so what is retired stays below 2x what is live
    if (w.ptr) g_split_ws_retired.push_back(w.ptr);
    w.ptr = nullptr; w.cap = 0;
    const size_t want = std::max(bytes + bytes / 2, (size_t)4 << 20);
    FA_CUDA(cudaMalloc(&w.ptr, want));
    w.cap = want;
  }
  *out = w.ptr;
  return FA_OK;
}

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// 4-D map over [batch, heads, n, d] with element strides (d contiguous); box = [128 bytes of d] x [box_rows rows] (128; the
// streamed tiles of the backward kernel: 64).
// Encoded maps are kept in a small per-thread cache keyed by everything that goes into them: a repeated call on the
// same tensors (the common case in a serving / benchmark loop) skips the four driver encodes.
struct MapKey {
  const void* ptr; int elem_size; int dt; int swizzle; int d; int64_t batch, heads, n, sb, sh, sn; int box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && elem_size == o.elem_size && dt == o.dt && swizzle == o.swizzle && d == o.d && batch == o.batch &&
           heads == o.heads && n == o.n && sb == o.sb && sh == o.sh && sn == o.sn && box_rows == o.box_rows;
  }
};
constexpr int kMapCacheSize = 32;
struct MapCache { MapKey key[kMapCacheSize]; CUtensorMap map[kMapCacheSize]; int used = 0; int next = 0; };
thread_local MapCache t_map_cache;

// `elem` is the fa_dtype of the tensor's elements.  The box is always 128 bytes of d wide: a head dim that does not fill
// its last box (or is smaller than one box) is zero-filled on load and clipped on store by TMA.
int make_map(CUtensorMap* out, const void* ptr, int elem_size, int elem, int64_t batch, int64_t heads, int64_t n, int d,
             int64_t sb, int64_t sh, int64_t sn, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, bool tf32_convert = false,
             int box_rows = 128) {
  CUtensorMapDataType dt = elem == FA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                         : elem == FA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (elem == FA_F32 && tf32_convert) dt = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;  // TMA converts fp32 -> tf32 while loading
  const MapKey key{ptr, elem_size, (int)dt, (int)swizzle, d, batch, heads, n, sb, sh, sn, box_rows};
  MapCache& mc = t_map_cache;
  for (int i = 0; i < mc.used; ++i)
    if (mc.key[i] == key) {
      *out = mc.map[i];
      return FA_OK;
    }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    snprintf(t_cuda_err, sizeof(t_cuda_err), "cuTensorMapEncodeTiled entry point not available");
    return FA_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return FA_ERR_ALIGNMENT;
  if (((sn * elem_size) & 15) || ((sh * elem_size) & 15) || ((sb * elem_size) & 15)) return FA_ERR_ALIGNMENT;
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)n, (cuuint64_t)heads, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)(sn * elem_size), (cuuint64_t)(sh * elem_size), (cuuint64_t)(sb * elem_size)};
  // A size-1 axis may come with stride 0 (its coordinate is always 0, so any stride TMA accepts will do: it wants a positive
  // multiple of 16).  A broadcast axis (stride 0, size > 1: MQA/GQA-style expanded K/V) cannot be described by a tiled map.
  for (int i = 0; i < 3; ++i) {
    if (strides[i] != 0) continue;
    if (dims[i + 1] != 1) return FA_ERR_UNSUPPORTED;
    strides[i] = 16;
  }
  cuuint32_t box[4] = {(cuuint32_t)(128 / elem_size), (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(t_cuda_err, sizeof(t_cuda_err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FA_ERR_CUDA;
  }
  const int slot = mc.used < kMapCacheSize ? mc.used++ : (mc.next = (mc.next + 1) % kMapCacheSize);
  mc.key[slot] = key;
  mc.map[slot] = *out;
  return FA_OK;
}

template <bool kTF32, int kHeadDim, bool kCausal, bool kOutF32, bool kF16, bool kPrecise = false>
int launch_tc(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo, const fa::FwdParams& fp,
              cudaStream_t st) {
  using T = fa::FwdTraits<kTF32, kHeadDim, kOutF32, kPrecise>;
  auto kern = fa::fa_fwd_sm100_kernel<kTF32, kHeadDim, kCausal, kOutF32, kF16, kPrecise>;
  static bool attr_set[64] = {};  // per kernel instance and per device (the attribute is per device); benign race (idempotent)
  int dev = 0;
  FA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return FA_ERR_NO_DEVICE;
  if (!attr_set[dev]) {
    FA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kSmemBytes));
    attr_set[dev] = true;
  }
  // one persistent CTA per SM pulling items from the work counter (fp.work_counter != nullptr), or one CTA per item
  const int64_t grid = fp.work_counter ? std::min<int64_t>(fp.n_items, std::max(1, current_sm_count())) : (int64_t)fp.n_items;
  if (grid <= 0 || grid > 0x7fffffff) return FA_ERR_INVALID_ARG;
  kern<<<(unsigned)grid, fa::kNumThreads, T::kSmemBytes, st>>>(mq, mk, mv, mo, fp);
  FA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

// ---- backward (fa_bwd_sm100.cuh): statistics pass + the dK/dV launch + the dQ launch ----
template <int kHeadDim, bool kF16, bool kDKV>
int launch_bwd(const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& t1, const CUtensorMap& t2, const fa::BwdParams& bp,
               dim3 grid, cudaStream_t st) {
  using T = fa::BwdTraits<kHeadDim>;
  auto kern = fa::fa_bwd_sm100_kernel<kHeadDim, kF16, kDKV>;
  static bool attr_set[64] = {};
  int dev = 0;
  FA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return FA_ERR_NO_DEVICE;
  if (!attr_set[dev]) {
    FA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kSmemBytes));
    attr_set[dev] = true;
  }
  kern<<<grid, fa::kBwdThreads, T::kSmemBytes, st>>>(r1, r2, t1, t2, bp);
  FA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

template <bool kDKV>
int launch_bwd_d(int di, bool f16, const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& t1, const CUtensorMap& t2,
                 const fa::BwdParams& bp, dim3 grid, cudaStream_t st) {
  if (di == 64) return f16 ? launch_bwd<64, true, kDKV>(r1, r2, t1, t2, bp, grid, st) : launch_bwd<64, false, kDKV>(r1, r2, t1, t2, bp, grid, st);
  if (di == 128) return f16 ? launch_bwd<128, true, kDKV>(r1, r2, t1, t2, bp, grid, st) : launch_bwd<128, false, kDKV>(r1, r2, t1, t2, bp, grid, st);
  return FA_ERR_UNSUPPORTED;
}

int run_bwd(const fa_bwd_params* p, cudaStream_t st) {
  const int dt = p->dtype;
  const bool f16 = dt == FA_F16;
  const int di = p->head_dim <= 64 ? 64 : 128;
  const int64_t kv_heads = p->kv_heads > 0 ? p->kv_heads : p->heads;
  const int64_t n_q_pad = (p->n_q + 127) / 128 * 128;
  const int64_t rows_pad = p->batch * p->heads * n_q_pad;
  int rc = FA_OK;
  void* ws = nullptr;
  void* ws_async = nullptr;
  if ((rc = split_workspace(st, (size_t)rows_pad * 2 * sizeof(float), &ws, &ws_async))) return rc;
  struct WsFree {
    void* p; cudaStream_t st;
    ~WsFree() { if (p) cudaFreeAsync(p, st); }
  } ws_free{ws_async, st};
  float* l2 = static_cast<float*>(ws);
  float* dsum = l2 + rows_pad;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap mq, mg, mk, mv;   // one box shape (128 rows x 128 bytes) serves the resident and the streamed role of each tensor
  if ((rc = make_map(&mq, p->q, 2, dt, p->batch, p->heads, p->n_q, p->head_dim, p->q_stride_b, p->q_stride_h, p->q_stride_n, sw))) return rc;
  if ((rc = make_map(&mg, p->d_o, 2, dt, p->batch, p->heads, p->n_q, p->head_dim, p->do_stride_b, p->do_stride_h, p->do_stride_n, sw))) return rc;
  if ((rc = make_map(&mk, p->k, 2, dt, p->batch, kv_heads, p->n_k, p->head_dim, p->k_stride_b, p->k_stride_h, p->k_stride_n, sw))) return rc;
  if ((rc = make_map(&mv, p->v, 2, dt, p->batch, kv_heads, p->n_k, p->head_dim, p->v_stride_b, p->v_stride_h, p->v_stride_n, sw))) return rc;
  // statistics pass (after the tensor maps: building them checks the alignment of every operand, dO included)
  {
    int tpr = 1;                                       // threads per row: 16-byte pieces, rounded up to a power of two
    while (tpr * 8 < p->head_dim) tpr *= 2;
    const int64_t blocks = (rows_pad * tpr + 255) / 256;
    if (blocks > 0x7fffffff) return FA_ERR_INVALID_ARG;
    if (f16)
      fa::fa_bwd_prep_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(
          static_cast<const __half*>(p->o), p->o_stride_b, p->o_stride_h, p->o_stride_n, static_cast<const __half*>(p->d_o), p->do_stride_b,
          p->do_stride_h, p->do_stride_n, p->lse, l2, dsum, (int)p->heads, (int)p->n_q, (int)n_q_pad, p->head_dim, tpr, rows_pad);
    else
      fa::fa_bwd_prep_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(p->o), p->o_stride_b, p->o_stride_h, p->o_stride_n, static_cast<const __nv_bfloat16*>(p->d_o),
          p->do_stride_b, p->do_stride_h, p->do_stride_n, p->lse, l2, dsum, (int)p->heads, (int)p->n_q, (int)n_q_pad, p->head_dim, tpr, rows_pad);
    FA_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  fa::BwdParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.scale = p->scale;
  bp.scale_log2 = p->scale * 1.4426950408889634f;
  bp.n_q = (int)p->n_q; bp.n_k = (int)p->n_k; bp.heads = (int)p->heads; bp.kv_heads = (int)kv_heads; bp.batch = (int)p->batch;
  bp.kv_group = (int)(p->heads / kv_heads);
  bp.head_dim = p->head_dim;
  bp.causal = p->causal != 0;
  bp.causal_offset = (int)(p->n_k - p->n_q);
  bp.n_q_pad = (int)n_q_pad;
  bp.l2 = l2; bp.dsum = dsum;
  bp.trace = nullptr;
#if FA_BWD_TRACE
  // tracing build: FA_B200_BWD_TRACE=path dumps the timeline of CTA (0,0,0) of the dK/dV launch (FA_B200_BWD_TRACE_DQ=1: of the dQ launch)
  static unsigned long long* d_btrace = nullptr;
  const char* btrace_path = getenv("FA_B200_BWD_TRACE");
  const bool btrace_dq = getenv("FA_B200_BWD_TRACE_DQ") != nullptr;
  const size_t btrace_n = 4 * 32 * 8;
  if (btrace_path) {
    if (!d_btrace) FA_CUDA(cudaMalloc(&d_btrace, btrace_n * sizeof(unsigned long long)));
    FA_CUDA(cudaMemsetAsync(d_btrace, 0, btrace_n * sizeof(unsigned long long), st));
  }
  struct BTraceDump {
    const char* path; unsigned long long* dev; size_t n; cudaStream_t st;
    ~BTraceDump() {
      if (!path) return;
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> h(n);
      cudaMemcpy(h.data(), dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      if (FILE* f = fopen(path, "w")) {
        for (size_t i = 0; i < n; ++i) fprintf(f, "%llu%c", h[i], (i % 8 == 7) ? '\n' : ' ');
        fclose(f);
      }
    }
  } btrace_dump{btrace_path, d_btrace, btrace_n, st};
#endif
  {   // dV, dK: one CTA per 128 keys of a K/V head
    fa::BwdParams b = bp;
    b.out0 = p->dv; b.o0_sb = p->dv_stride_b; b.o0_sh = p->dv_stride_h; b.o0_sn = p->dv_stride_n;
    b.out1 = p->dk; b.o1_sb = p->dk_stride_b; b.o1_sh = p->dk_stride_h; b.o1_sn = p->dk_stride_n;
#if FA_BWD_TRACE
    if (btrace_path && !btrace_dq) b.trace = d_btrace;
#endif
    const dim3 grid((unsigned)((p->n_k + 127) / 128), (unsigned)kv_heads, (unsigned)p->batch);
    if ((rc = launch_bwd_d<true>(di, f16, mk, mv, mq, mg, b, grid, st))) return rc;
  }
  {   // dQ: one CTA per 128 query rows of a head
    fa::BwdParams b = bp;
    b.out0 = p->dq; b.o0_sb = p->dq_stride_b; b.o0_sh = p->dq_stride_h; b.o0_sn = p->dq_stride_n;
#if FA_BWD_TRACE
    if (btrace_path && btrace_dq) b.trace = d_btrace;
#endif
    const dim3 grid((unsigned)((p->n_q + 127) / 128), (unsigned)p->heads, (unsigned)p->batch);
    if ((rc = launch_bwd_d<false>(di, f16, mq, mg, mk, mv, b, grid, st))) return rc;
  }
  return FA_OK;
}

template <bool kTF32, int kHeadDim, bool kOutF32, bool kF16 = false>
int launch_tc_c(bool causal, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
                const fa::FwdParams& fp, cudaStream_t st) {
  return causal ? launch_tc<kTF32, kHeadDim, true, kOutF32, kF16>(mq, mk, mv, mo, fp, st)
                : launch_tc<kTF32, kHeadDim, false, kOutF32, kF16>(mq, mk, mv, mo, fp, st);
}

// The tcgen05 kernel is instantiated for tile rows of 128, 256 and 512 bytes: fp32 head dims 32 / 64 / 128, 16-bit head
// dims 64 / 128 / 256 (the 512-byte instances keep one Q tile per CTA: FwdTraits::kSlots).  Any smaller head dim whose
// rows keep TMA's 16-byte alignment runs on the next instance up: TMA zero-fills the missing columns of Q, K and V in SMEM
// (zero columns add nothing to q.k, and give zero O columns) and clips them from the O store.
// 0 = no instance (fp32 d > 128: CUDA-core kernel; FA_B200_NO_WIDE=1 also sends the 512-byte rows there, A/B aid).
// FA_FLAG_PRECISE (or FA_B200_PRECISE=1 in the environment) on fp32 inputs: the 3xTF32 instance (head dim <= 64), else the
// fp32 CUDA-core kernel.  16-bit inputs have no precise mode: their operands are exact already.
bool want_precise(const fa_params* p) {
  static const bool env = [] { const char* e = getenv("FA_B200_PRECISE"); return e && atoi(e) != 0; }();
  return p->dtype == FA_F32 && (env || (p->flags & FA_FLAG_PRECISE));
}

int tc_instance_dim(const fa_params* p) {
  const int d = p->head_dim;
  if (want_precise(p)) return (d % 4 == 0 && d <= 64) ? 64 : 0;
  static const bool no_wide = [] { const char* e = getenv("FA_B200_NO_WIDE"); return e && atoi(e) != 0; }();
  int di;
  if (p->dtype == FA_F32) di = (d % 4 || d > 128) ? 0 : (d <= 32 ? 32 : (d <= 64 ? 64 : 128));
  else di = (d % 8 || d > 256) ? 0 : (d <= 64 ? 64 : (d <= 128 ? 128 : 256));
  if (no_wide && di * (p->dtype == FA_F32 ? 4 : 2) > 256) di = 0;
  return di;
}
bool tc_supported(const fa_params* p) { return tc_instance_dim(p) != 0; }

int run_tc(const fa_params* p, cudaStream_t st) {
  const bool bf16 = p->dtype != FA_F32;   // 16-bit operands (bf16 or fp16): kind::f16 instances
  const int in_dt = p->dtype;
  const int in_sz = bf16 ? 2 : 4;
  const bool precise = want_precise(p);
  // ---- split-KV across CTAs (flash-decoding): a launch with far fewer 128-row Q tiles than SMs and a long key sequence cuts
  // every Q tile's K/V tiles into `kv_splits` runs, one item each, and merges the runs' partials afterwards.  The reference
  // sizes its grid by (batch, ceil(N / 32)) only (src/flashattention.cu:592) and leaves such launches on a handful of SMs.
  // Off for accumulate mode and FA_FLAG_BATCH_INVARIANT (the split depends on the launch size); FA_B200_KV_SPLIT=0 / n: off / force.
  int kv_splits = 1, kv_chunk_tiles = 0;
  {
    const int64_t kv_tiles = (p->n_k + fa::kBlockN - 1) / fa::kBlockN;
    const int64_t q_tiles = p->batch * p->heads * ((p->n_q + fa::kBlockM - 1) / fa::kBlockM);
    const int64_t sms = std::max(1, current_sm_count());
    int64_t want = 1;
    if (const char* e = getenv("FA_B200_KV_SPLIT")) want = std::max(1, atoi(e));
    else if (2 * q_tiles <= sms && kv_tiles >= 8) want = std::min<int64_t>(std::min<int64_t>(sms / q_tiles, kv_tiles / 4), 64);
    if (want > 1 && kv_tiles > 1 && !p->o_acc && !(p->flags & FA_FLAG_BATCH_INVARIANT)) {
      kv_chunk_tiles = (int)((kv_tiles + want - 1) / want);
      kv_splits = (int)((kv_tiles + kv_chunk_tiles - 1) / kv_chunk_tiles);
      if (kv_splits <= 1) { kv_splits = 1; kv_chunk_tiles = 0; }
    }
  }
  const bool out_f32 = !bf16 || p->o_f32 || kv_splits > 1;   // the runs' partials are fp32 whatever the caller's O is
  const int out_sz = out_f32 ? 4 : 2;
  CUtensorMap mq, mk, mv, mo;
  int rc = FA_OK;
  // workspace of the runs' partials: O [kv_splits * batch, heads, n_q, d] fp32 + LSE [kv_splits * batch, heads, n_q], stream-ordered
  float* ws_o = nullptr;
  float* ws_lse = nullptr;
  const int64_t ws_rows = p->batch * p->heads * p->n_q;
  void* ws_async = nullptr;     // a stream-ordered allocation of this call (captured launches only), freed when run_tc returns
  if (kv_splits > 1) {
    void* ws = nullptr;
    if ((rc = split_workspace(st, (size_t)kv_splits * ws_rows * (p->head_dim + 1) * sizeof(float), &ws, &ws_async))) return rc;
    ws_o = static_cast<float*>(ws);
    ws_lse = ws_o + (size_t)kv_splits * ws_rows * p->head_dim;
  }
  struct WsFree {
    void* p; cudaStream_t st;
    ~WsFree() { if (p) cudaFreeAsync(p, st); }
  } ws_free{ws_async, st};
  // fp32 tensors are loaded through TFLOAT32 tensor maps: TMA rounds fp32 -> tf32 to nearest on the way into SMEM, which
  // removes the truncation bias the tensor core would otherwise apply (measured on B200: max error 4.1e-4 -> 9.8e-5 on C1).
  static const bool tf32_tma_env = [] { const char* e = getenv("FA_B200_TMA_TF32"); return !(e && atoi(e) == 0); }();
  const bool tf32_tma = tf32_tma_env && !precise;   // a precise instance needs the fp32 bits as they are (hi = trunc, lo = rest)
  if ((rc = make_map(&mq, p->q, in_sz, in_dt, p->batch, p->heads, p->n_q, p->head_dim, p->q_stride_b, p->q_stride_h, p->q_stride_n, CU_TENSOR_MAP_SWIZZLE_128B, tf32_tma))) return rc;
  const int64_t kv_heads = p->kv_heads > 0 ? p->kv_heads : p->heads;
  if ((rc = make_map(&mk, p->k, in_sz, in_dt, p->batch, kv_heads, p->n_k, p->head_dim, p->k_stride_b, p->k_stride_h, p->k_stride_n, CU_TENSOR_MAP_SWIZZLE_128B, tf32_tma))) return rc;
  // V is the MN-major B operand of P*V: bf16 uses the ordinary 128B swizzle; 32-bit (tf32) MN-major operands must
  // be in the SWIZZLE_128B_BASE32B layout (32-byte units over 4-row groups), written by TMA's 128B_ATOM_32B mode.
  uint32_t v_lbo = fa::kChunkBytes, v_sbo = bf16 ? 1024 : 512, v_layout = bf16 ? fa::kLayoutSw128 : fa::kLayoutSw128Base32;
  CUtensorMapSwizzle v_swz = bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (const char* dbg = getenv("FA_B200_V_VARIANT")) {  // bring-up aid: alternative descriptor hypotheses for tf32
    const int vv = atoi(dbg);
    if (!bf16 && vv == 2) v_sbo = 1024;
    if (!bf16 && vv == 3) { v_lbo = 512; v_sbo = fa::kChunkBytes; }
    if (!bf16 && vv == 4) { v_layout = fa::kLayoutSw128; v_sbo = 1024; v_swz = CU_TENSOR_MAP_SWIZZLE_128B; }
  }
  if ((rc = make_map(&mv, p->v, in_sz, in_dt, p->batch, kv_heads, p->n_k, p->head_dim, p->v_stride_b, p->v_stride_h, p->v_stride_n, v_swz, tf32_tma))) return rc;
  if (kv_splits > 1) {
    const int64_t hd = p->head_dim;
    if ((rc = make_map(&mo, ws_o, 4, (int)FA_F32, (int64_t)kv_splits * p->batch, p->heads, p->n_q, p->head_dim, p->heads * p->n_q * hd, p->n_q * hd, hd))) return rc;
  } else if ((rc = make_map(&mo, p->o, out_sz, out_f32 ? (int)FA_F32 : in_dt, p->batch, p->heads, p->n_q, p->head_dim, p->o_stride_b, p->o_stride_h, p->o_stride_n))) return rc;
  fa::FwdParams fp;
  fp.scale = p->scale;
  fp.scale_log2 = p->scale * 1.4426950408889634f;
  fp.n_q = (int)p->n_q; fp.n_k = (int)p->n_k; fp.heads = (int)p->heads; fp.batch = (int)p->batch;
  fp.causal_offset = (int)(p->n_k - p->n_q);
  fp.num_m_blocks = (int)((p->n_q + 2 * fa::kBlockM - 1) / (2 * fa::kBlockM));
  fp.lse = p->lse;
  fp.o_ptr = p->o;
  fp.o_sb = p->o_stride_b; fp.o_sh = p->o_stride_h; fp.o_sn = p->o_stride_n;
  fp.o_row_bytes = p->head_dim * out_sz;
  fp.kv_splits = kv_splits; fp.kv_chunk_tiles = kv_chunk_tiles;
  fp.kv_group = (int)(p->heads / kv_heads);
  if (kv_splits > 1) {
    fp.lse = ws_lse;
    fp.o_ptr = ws_o;
    fp.o_sn = p->head_dim; fp.o_sh = p->n_q * (int64_t)p->head_dim; fp.o_sb = p->heads * fp.o_sh;
  }
  fp.acc_o = p->o_acc; fp.acc_lse = p->lse_acc; fp.head_dim = p->head_dim;
  {
    // whole waves of 256-row CTAs; a remainder of at most SMs/2 blocks runs as twice as many 128-row CTAs (one wave)
    const int64_t nb = (int64_t)fp.num_m_blocks * fp.heads * fp.batch * kv_splits;
    const int64_t sms = std::max(1, current_sm_count());
    int64_t n_big = (nb / sms) * sms;
    if (2 * (nb - n_big) > sms || getenv("FA_B200_NO_SPLIT_WAVE")) n_big = nb;
    const bool one_slot = tc_instance_dim(p) * in_sz > 256 || precise;   // 512-byte rows (or 256 + their lo copy): one Q tile per CTA, every item is a 128-row item
    if (one_slot) n_big = 0;
    if (n_big > 0x3fffffff) return FA_ERR_INVALID_ARG;
    fp.n_big = (int)n_big;
    // the remainder CTAs split their K/V range over the two tile slots (FA_B200_TAIL_SPLIT=0: one slot, A/B aid)
    const char* ts = getenv("FA_B200_TAIL_SPLIT");   // read per call so a test can compare both modes in one process
    fp.tail_split = ((ts && atoi(ts) == 0) || (p->flags & FA_FLAG_BATCH_INVARIANT) || one_slot) ? 0 : 1;
    const int64_t n_items = n_big + 2 * (nb - n_big);
    if (n_items > 0x7fffffff) return FA_ERR_INVALID_ARG;
    fp.n_items = (int)n_items;
    // persistent CTAs take items from a device counter (FA_B200_PERSISTENT=0: one CTA per item, hardware-scheduled)
    const char* ps = getenv("FA_B200_PERSISTENT");
    fp.work_counter = nullptr;
    if (!(ps && atoi(ps) == 0)) {
      if ((rc = next_work_counter(st, &fp.work_counter))) return rc;
    }
  }
  fp.v_desc_hi = fa::make_sdesc_hi(v_lbo, v_sbo, v_layout);
  fp.trace = nullptr;
#if FA_TRACE
  // tracing build: record CTA 0's pipeline timeline and dump it (synchronously) to $FA_B200_TRACE after the launch
  static unsigned long long* d_trace = nullptr;
  const char* trace_path = getenv("FA_B200_TRACE");
  const size_t trace_n = 4 * fa::kTraceSteps * 8;
  if (trace_path) {
    if (!d_trace) FA_CUDA(cudaMalloc(&d_trace, trace_n * sizeof(unsigned long long)));
    FA_CUDA(cudaMemsetAsync(d_trace, 0, trace_n * sizeof(unsigned long long), st));
    fp.trace = d_trace;
  }
  struct TraceDump {
    const char* path; unsigned long long* dev; size_t n; cudaStream_t st;
    ~TraceDump() {
      if (!path) return;
      cudaStreamSynchronize(st);
      unsigned long long* h = (unsigned long long*)malloc(n * sizeof(unsigned long long));
      cudaMemcpy(h, dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      FILE* f = fopen(path, "w");
      if (f) {
        for (size_t i = 0; i < n; ++i) fprintf(f, "%llu%c", h[i], (i % 8 == 7) ? '\n' : ' ');
        fclose(f);
      }
      free(h);
    }
  } trace_dump{trace_path, d_trace, trace_n, st};
#endif
  const bool c = p->causal != 0;
  const int di = tc_instance_dim(p);   // kernel instance (>= head_dim; the tensor maps carry the true head dim)
  const bool f16 = p->dtype == FA_F16;
  auto launch = [&]() -> int {
  if (precise) {
    return c ? launch_tc<true, 64, true, false, false, true>(mq, mk, mv, mo, fp, st)
             : launch_tc<true, 64, false, false, false, true>(mq, mk, mv, mo, fp, st);
  }
  if (!bf16) {
    if (di == 32) return launch_tc_c<true, 32, false>(c, mq, mk, mv, mo, fp, st);
    if (di == 64) return launch_tc_c<true, 64, false>(c, mq, mk, mv, mo, fp, st);
    if (di == 128) return launch_tc_c<true, 128, false>(c, mq, mk, mv, mo, fp, st);
  } else if (!out_f32) {
    if (di == 64) return f16 ? launch_tc_c<false, 64, false, true>(c, mq, mk, mv, mo, fp, st) : launch_tc_c<false, 64, false>(c, mq, mk, mv, mo, fp, st);
    if (di == 128) return f16 ? launch_tc_c<false, 128, false, true>(c, mq, mk, mv, mo, fp, st) : launch_tc_c<false, 128, false>(c, mq, mk, mv, mo, fp, st);
    if (di == 256) return f16 ? launch_tc_c<false, 256, false, true>(c, mq, mk, mv, mo, fp, st) : launch_tc_c<false, 256, false>(c, mq, mk, mv, mo, fp, st);
  } else {
    if (di == 64) return f16 ? launch_tc_c<false, 64, true, true>(c, mq, mk, mv, mo, fp, st) : launch_tc_c<false, 64, true>(c, mq, mk, mv, mo, fp, st);
    if (di == 128) return f16 ? launch_tc_c<false, 128, true, true>(c, mq, mk, mv, mo, fp, st) : launch_tc_c<false, 128, true>(c, mq, mk, mv, mo, fp, st);
    if (di == 256) return f16 ? launch_tc_c<false, 256, true, true>(c, mq, mk, mv, mo, fp, st) : launch_tc_c<false, 256, true>(c, mq, mk, mv, mo, fp, st);
  }
  return FA_ERR_UNSUPPORTED;
  };
  if ((rc = launch())) return rc;
  if (kv_splits > 1) {
    // merge the runs' partials into the caller's O (dtype, strides) and LSE
    const unsigned grid = (unsigned)((ws_rows + 7) / 8);
    const int H = (int)p->heads, nq = (int)p->n_q, hd = p->head_dim;
    if (!bf16 || p->o_f32)
      fa::fa_combine_splits_kernel<float><<<grid, 256, 0, st>>>(ws_o, ws_lse, kv_splits, ws_rows, H, nq, hd, static_cast<float*>(p->o),
                                                                 p->o_stride_b, p->o_stride_h, p->o_stride_n, p->lse);
    else if (f16)
      fa::fa_combine_splits_kernel<__half><<<grid, 256, 0, st>>>(ws_o, ws_lse, kv_splits, ws_rows, H, nq, hd, static_cast<__half*>(p->o),
                                                                  p->o_stride_b, p->o_stride_h, p->o_stride_n, p->lse);
    else
      fa::fa_combine_splits_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(ws_o, ws_lse, kv_splits, ws_rows, H, nq, hd,
                                                                         static_cast<__nv_bfloat16*>(p->o), p->o_stride_b, p->o_stride_h,
                                                                         p->o_stride_n, p->lse);
    FA_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return FA_OK;
}

template <typename TIn, typename TOut>
int launch_simt(const fa::SimtParams& sp, cudaStream_t st) {
  const size_t smem = sizeof(float) * (fa::kSimtKeys * (sp.head_dim + 1) + fa::kSimtKeys * sp.head_dim + fa::kSimtRows * sp.head_dim);
  auto kern = fa::fa_fwd_simt_kernel<TIn, TOut>;
  if (smem > 48 * 1024) FA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((sp.n_q + fa::kSimtRows - 1) / fa::kSimtRows), (unsigned)(sp.batch * sp.heads));
  kern<<<grid, fa::kSimtRows * 32, smem, st>>>(sp);
  FA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

int run_simt(const fa_params* p, cudaStream_t st) {
  if (p->head_dim > fa::kSimtMaxD || p->head_dim % 8 != 0) return FA_ERR_UNSUPPORTED;
  if (p->batch * p->heads > 65535) return FA_ERR_UNSUPPORTED;
  fa::SimtParams sp;
  sp.q = p->q; sp.k = p->k; sp.v = p->v; sp.o = p->o; sp.lse = p->lse;
  sp.q_sb = p->q_stride_b; sp.q_sh = p->q_stride_h; sp.q_sn = p->q_stride_n;
  sp.k_sb = p->k_stride_b; sp.k_sh = p->k_stride_h; sp.k_sn = p->k_stride_n;
  sp.v_sb = p->v_stride_b; sp.v_sh = p->v_stride_h; sp.v_sn = p->v_stride_n;
  sp.o_sb = p->o_stride_b; sp.o_sh = p->o_stride_h; sp.o_sn = p->o_stride_n;
  sp.n_q = (int)p->n_q; sp.n_k = (int)p->n_k; sp.heads = (int)p->heads; sp.batch = (int)p->batch; sp.head_dim = p->head_dim;
  sp.causal = p->causal; sp.causal_offset = (int)(p->n_k - p->n_q); sp.scale = p->scale;
  sp.kv_group = (int)(p->heads / (p->kv_heads > 0 ? p->kv_heads : p->heads));
  if (p->dtype == FA_F32) return launch_simt<float, float>(sp, st);
  if (p->dtype == FA_F16) return p->o_f32 ? launch_simt<__half, float>(sp, st) : launch_simt<__half, __half>(sp, st);
  if (p->o_f32) return launch_simt<__nv_bfloat16, float>(sp, st);
  return launch_simt<__nv_bfloat16, __nv_bfloat16>(sp, st);
}

void fill_contiguous(fa_params* p, const void* q, const void* k, const void* v, void* o, float* lse, int64_t batch, int64_t heads,
                     int64_t n_q, int64_t n_k, int32_t d, float scale, int32_t causal, int32_t dtype) {
  memset(p, 0, sizeof(*p));
  p->q = q; p->k = k; p->v = v; p->o = o; p->lse = lse;
  p->batch = batch; p->heads = heads; p->n_q = n_q; p->n_k = n_k; p->head_dim = d; p->dtype = dtype; p->causal = causal;
  p->scale = scale;
  p->q_stride_n = d; p->q_stride_h = n_q * d; p->q_stride_b = heads * n_q * d;
  p->k_stride_n = d; p->k_stride_h = n_k * d; p->k_stride_b = heads * n_k * d;
  p->v_stride_n = d; p->v_stride_h = n_k * d; p->v_stride_b = heads * n_k * d;
  p->o_stride_n = d; p->o_stride_h = n_q * d; p->o_stride_b = heads * n_q * d;
}

// cached device scratch + streams for fa_forward_host
constexpr int kHostMaxChunks = 16;
struct HostScratch {
  void* q = nullptr; void* k = nullptr; void* v = nullptr; void* o = nullptr;
  size_t cap_q = 0, cap_kv = 0, cap_o = 0;
  cudaStream_t st_in = nullptr, st_run = nullptr, st_out = nullptr;
  cudaEvent_t ev_in[kHostMaxChunks] = {}, ev_run[kHostMaxChunks] = {};
};
// one per device: the buffers, streams and events belong to the device that was current when they were made
HostScratch g_hs[64];
std::mutex g_hs_mu[64];

}  // namespace

extern "C" {

int fa_version(void) { return FA_B200_VERSION; }
const char* fa_last_cuda_error(void) { return t_cuda_err; }
int fa_last_impl(void) { return t_last_impl; }
int64_t fa_launch_count(void) { return g_launches.load(); }
int fa_watchdog_info(uint32_t out[4]) {
  if (!out) return FA_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) out[i] = g_wd_host ? g_wd_host[i] : 0u;
  return FA_OK;
}

const char* fa_strerror(int status) {
  switch (status) {
    case FA_OK: return "ok";
    case FA_ERR_INVALID_ARG: return "invalid argument";
    case FA_ERR_NO_DEVICE: return "no usable CUDA device (this library needs an sm_100 GPU; there is no CPU fallback)";
    case FA_ERR_CUDA: return "CUDA call failed (see fa_last_cuda_error)";
    case FA_ERR_UNSUPPORTED: return "unsupported shape/dtype combination";
    case FA_ERR_ALIGNMENT: return "pointer or stride is not 16-byte aligned";
    default: return "unknown status";
  }
}

int fa_query_instance(int32_t dtype, int32_t head_dim) {
  if ((dtype != FA_F32 && dtype != FA_BF16 && dtype != FA_F16) || head_dim <= 0) return FA_ERR_INVALID_ARG;
  fa_params p;
  memset(&p, 0, sizeof(p));
  p.dtype = dtype;
  p.head_dim = head_dim;
  const int di = tc_instance_dim(&p);
  if (di) return di;
  return (head_dim <= fa::kSimtMaxD && head_dim % 8 == 0) ? 0 : FA_ERR_UNSUPPORTED;
}

int fa_forward_ex(const fa_params* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->o) return FA_ERR_INVALID_ARG;
  if (p->batch <= 0 || p->heads <= 0 || p->n_q <= 0 || p->n_k <= 0 || p->head_dim <= 0) return FA_ERR_INVALID_ARG;
  if (p->dtype != FA_F32 && p->dtype != FA_BF16 && p->dtype != FA_F16) return FA_ERR_INVALID_ARG;
  if (!(p->scale > 0.f) || !std::isfinite(p->scale)) return FA_ERR_INVALID_ARG;
  if (p->n_q > 0x7fffffff || p->n_k > 0x7fffffff || p->batch > 0x7fffffff || p->heads > 0x7fffffff) return FA_ERR_INVALID_ARG;
  if (p->o_f32 && p->dtype == FA_F32) return FA_ERR_INVALID_ARG;   // fp32 inputs already give fp32 O
  if ((p->o_acc == nullptr) != (p->lse_acc == nullptr)) return FA_ERR_INVALID_ARG;
  if (p->kv_heads < 0 || (p->kv_heads > 0 && (p->kv_heads > p->heads || p->heads % p->kv_heads != 0))) return FA_ERR_INVALID_ARG;
  if (p->o_acc && ((reinterpret_cast<uintptr_t>(p->o_acc) & 15) || p->head_dim % 4)) return FA_ERR_ALIGNMENT;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  if (major != 10) return FA_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int impl = p->impl;
  if (impl == 0) impl = tc_supported(p) ? FA_IMPL_TCGEN05 : FA_IMPL_SIMT;
  if (impl == FA_IMPL_TCGEN05) {
    if (!tc_supported(p)) return FA_ERR_UNSUPPORTED;
    rc = run_tc(p, st);
  } else if (impl == FA_IMPL_SIMT) {
    if (p->o_acc) return FA_ERR_UNSUPPORTED;   // accumulate mode lives in the tcgen05 kernel's epilogue
    rc = run_simt(p, st);
  } else {
    return FA_ERR_INVALID_ARG;
  }
  if (rc == FA_OK) t_last_impl = impl;
  return rc;
}

int fa_backward(const fa_bwd_params* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->o || !p->d_o || !p->lse || !p->dq || !p->dk || !p->dv) return FA_ERR_INVALID_ARG;
  if (p->batch <= 0 || p->heads <= 0 || p->n_q <= 0 || p->n_k <= 0 || p->head_dim <= 0) return FA_ERR_INVALID_ARG;
  if (p->n_q > 0x3fffffff || p->n_k > 0x3fffffff || p->batch > 65535 || p->heads > 65535) return FA_ERR_INVALID_ARG;
  if (!(p->scale > 0.f) || !std::isfinite(p->scale)) return FA_ERR_INVALID_ARG;
  if (p->kv_heads < 0 || (p->kv_heads > 0 && (p->kv_heads > p->heads || p->heads % p->kv_heads != 0))) return FA_ERR_INVALID_ARG;
  if (p->dtype != FA_BF16 && p->dtype != FA_F16) return p->dtype == FA_F32 ? FA_ERR_UNSUPPORTED : FA_ERR_INVALID_ARG;
  if (p->head_dim % 8 != 0 || p->head_dim > 128) return FA_ERR_UNSUPPORTED;
  // the epilogue writes 16-byte vectors and the statistics pass reads O with them (dO is checked with its tensor map)
  if ((reinterpret_cast<uintptr_t>(p->o) & 15) || p->o_stride_b % 8 || p->o_stride_h % 8 || p->o_stride_n % 8) return FA_ERR_ALIGNMENT;
  const void* outs[3] = {p->dq, p->dk, p->dv};
  const int64_t ostr[9] = {p->dq_stride_b, p->dq_stride_h, p->dq_stride_n, p->dk_stride_b, p->dk_stride_h, p->dk_stride_n,
                           p->dv_stride_b, p->dv_stride_h, p->dv_stride_n};
  for (int i = 0; i < 3; ++i)
    if (reinterpret_cast<uintptr_t>(outs[i]) & 15) return FA_ERR_ALIGNMENT;
  for (int i = 0; i < 9; ++i)
    if (ostr[i] % 8) return FA_ERR_ALIGNMENT;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  if (major != 10) return FA_ERR_NO_DEVICE;
  rc = run_bwd(p, static_cast<cudaStream_t>(stream));
  if (rc == FA_OK) t_last_impl = FA_IMPL_TCGEN05;
  return rc;
}

int fa_forward(const void* q, const void* k, const void* v, void* o, float* lse, int64_t batch, int64_t heads, int64_t n_q,
               int64_t n_k, int32_t head_dim, float scale, int32_t causal, int32_t dtype, void* stream) {
  fa_params p;
  fill_contiguous(&p, q, k, v, o, lse, batch, heads, n_q, n_k, head_dim, scale, causal, dtype);
  return fa_forward_ex(&p, stream);
}

int fa_forward_packed_qkv(const float* inp, float* out, float* lse, int32_t B, int32_t T, int32_t NH, int32_t hs, float scale,
                          int32_t causal, void* stream) {
  return fa_forward_packed_qkv_ex(inp, out, lse, B, T, NH, hs, scale, causal, 0, stream);
}

int fa_forward_packed_qkv_ex(const float* inp, float* out, float* lse, int32_t B, int32_t T, int32_t NH, int32_t hs, float scale,
                             int32_t causal, int32_t flags, void* stream) {
  if (!inp || !out || B <= 0 || T <= 0 || NH <= 0 || hs <= 0) return FA_ERR_INVALID_ARG;
  const int64_t C = (int64_t)NH * hs;
  fa_params p;
  memset(&p, 0, sizeof(p));
  p.q = inp; p.k = inp + C; p.v = inp + 2 * C; p.o = out; p.lse = lse;
  p.batch = B; p.heads = NH; p.n_q = T; p.n_k = T; p.head_dim = hs; p.dtype = FA_F32; p.causal = causal; p.scale = scale;
  p.q_stride_n = p.k_stride_n = p.v_stride_n = 3 * C;
  p.q_stride_h = p.k_stride_h = p.v_stride_h = hs;
  p.q_stride_b = p.k_stride_b = p.v_stride_b = (int64_t)T * 3 * C;
  p.o_stride_n = C; p.o_stride_h = hs; p.o_stride_b = (int64_t)T * C;
  p.flags = flags;
  return fa_forward_ex(&p, stream);
}

int fa_forward_host(const void* qh, const void* kh, const void* vh, void* oh, int64_t batch, int64_t heads, int64_t n_q, int64_t n_k,
                    int32_t head_dim, float scale, int32_t causal, int32_t dtype) {
  if (!qh || !kh || !vh || !oh) return FA_ERR_INVALID_ARG;
  if (batch <= 0 || heads <= 0 || n_q <= 0 || n_k <= 0 || head_dim <= 0) return FA_ERR_INVALID_ARG;
  if (dtype != FA_F32 && dtype != FA_BF16 && dtype != FA_F16) return FA_ERR_INVALID_ARG;
  if (!(scale > 0.f) || !std::isfinite(scale)) return FA_ERR_INVALID_ARG;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  if (major != 10) return FA_ERR_NO_DEVICE;
  const size_t es = dtype == FA_F32 ? 4 : 2;
  const size_t bq = (size_t)batch * heads * n_q * head_dim * es, bkv = (size_t)batch * heads * n_k * head_dim * es;
  int dev = 0;
  FA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return FA_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(g_hs_mu[dev]);
  HostScratch& s = g_hs[dev];
  if (!s.st_in) {
    FA_CUDA(cudaStreamCreateWithFlags(&s.st_in, cudaStreamNonBlocking));
    FA_CUDA(cudaStreamCreateWithFlags(&s.st_run, cudaStreamNonBlocking));
    FA_CUDA(cudaStreamCreateWithFlags(&s.st_out, cudaStreamNonBlocking));
    for (int i = 0; i < kHostMaxChunks; ++i) {
      FA_CUDA(cudaEventCreateWithFlags(&s.ev_in[i], cudaEventDisableTiming));
      FA_CUDA(cudaEventCreateWithFlags(&s.ev_run[i], cudaEventDisableTiming));
    }
  }
  if (s.cap_q < bq) {
    if (s.q) cudaFree(s.q);
    s.q = nullptr; s.cap_q = 0;
    FA_CUDA(cudaMalloc(&s.q, bq));
    s.cap_q = bq;
  }
  if (s.cap_o < bq) {
    if (s.o) cudaFree(s.o);
    s.o = nullptr; s.cap_o = 0;
    FA_CUDA(cudaMalloc(&s.o, bq));
    s.cap_o = bq;
  }
  if (s.cap_kv < bkv) {
    if (s.k) cudaFree(s.k);
    if (s.v) cudaFree(s.v);
    s.k = s.v = nullptr; s.cap_kv = 0;
    FA_CUDA(cudaMalloc(&s.k, bkv));
    FA_CUDA(cudaMalloc(&s.v, bkv));
    s.cap_kv = bkv;
  }
  // Every (batch, head) is independent, so the copy-in, the kernel and the copy-out are pipelined over chunks of
  // the flattened batch*heads axis on three streams: H2D of chunk g+1 and D2H of chunk g-1 overlap the kernel of g.
  // The job is bound by the H2D copies (3/4 of the bytes), which run back to back; what is left to shorten is the tail
  // after the last H2D — the last chunk's kernel and D2H — so the chunks shrink geometrically towards the end
  // (each chunk = kDecay of what remains, at least kMinChunkBytes of traffic and at least one head).
  const int64_t bh = batch * heads;
  const size_t row_q = (size_t)n_q * head_dim * es, row_kv = (size_t)n_k * head_dim * es;
  int64_t sched[kHostMaxChunks];
  int chunks = 0;
  {
    double decay = 0.4;
    if (const char* e = getenv("FA_B200_HOST_DECAY")) decay = std::min(1.0, std::max(0.05, atof(e)));
    const size_t head_bytes = 2 * row_q + 2 * row_kv;
    const int64_t min_heads = std::max<int64_t>(1, (int64_t)(((size_t)2 << 20) + head_bytes - 1) / (int64_t)head_bytes);
    int64_t left = bh;
    while (left > 0) {
      int64_t take = chunks + 1 == kHostMaxChunks ? left : std::max<int64_t>(min_heads, (int64_t)std::ceil(left * decay));
      take = std::min(take, left);
      sched[chunks++] = take;
      left -= take;
    }
    if (const char* e = getenv("FA_B200_HOST_CHUNKS")) {   // A/B aid: N equal chunks
      const int64_t n = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(bh, kHostMaxChunks), atoi(e)));
      const int64_t per = (bh + n - 1) / n;
      chunks = 0;
      for (int64_t h0 = 0; h0 < bh; h0 += per) sched[chunks++] = std::min<int64_t>(per, bh - h0);
    }
    if (const char* e = getenv("FA_B200_HOST_SCHED")) {    // A/B aid: explicit heads per chunk, e.g. "6,5,3,1,1" (rest -> last)
      chunks = 0;
      int64_t left2 = bh;
      const char* c = e;
      while (*c && left2 > 0 && chunks < kHostMaxChunks - 1) {
        const int64_t v = std::max<int64_t>(1, std::min<int64_t>(left2, atoll(c)));
        sched[chunks++] = v;
        left2 -= v;
        while (*c && *c != ',') ++c;
        if (*c == ',') ++c;
      }
      if (left2 > 0) sched[chunks++] = left2;
    }
  }
  int64_t h0 = 0;
  for (int g = 0; g < chunks; ++g) {
    const int64_t nh = sched[g];
    char* dq = (char*)s.q + h0 * row_q; char* dk = (char*)s.k + h0 * row_kv; char* dv = (char*)s.v + h0 * row_kv;
    char* dout = (char*)s.o + h0 * row_q;
    FA_CUDA(cudaMemcpyAsync(dq, (const char*)qh + h0 * row_q, nh * row_q, cudaMemcpyHostToDevice, s.st_in));
    FA_CUDA(cudaMemcpyAsync(dk, (const char*)kh + h0 * row_kv, nh * row_kv, cudaMemcpyHostToDevice, s.st_in));
    FA_CUDA(cudaMemcpyAsync(dv, (const char*)vh + h0 * row_kv, nh * row_kv, cudaMemcpyHostToDevice, s.st_in));
    FA_CUDA(cudaEventRecord(s.ev_in[g], s.st_in));
    FA_CUDA(cudaStreamWaitEvent(s.st_run, s.ev_in[g], 0));
    rc = fa_forward(dq, dk, dv, dout, nullptr, 1, nh, n_q, n_k, head_dim, scale, causal, dtype, s.st_run);
    if (rc) return rc;
    FA_CUDA(cudaEventRecord(s.ev_run[g], s.st_run));
    FA_CUDA(cudaStreamWaitEvent(s.st_out, s.ev_run[g], 0));
    FA_CUDA(cudaMemcpyAsync((char*)oh + h0 * row_q, dout, nh * row_q, cudaMemcpyDeviceToHost, s.st_out));
    h0 += nh;
  }
  FA_CUDA(cudaStreamSynchronize(s.st_out));
  return FA_OK;
}

int fa_merge_partials(float* o_acc, float* lse_acc, const float* o_new, const float* lse_new, int64_t rows, int32_t head_dim,
                      void* stream) {
  if (!o_acc || !lse_acc || !o_new || !lse_new || rows <= 0 || head_dim <= 0 || head_dim % 4) return FA_ERR_INVALID_ARG;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = rows * (head_dim / 4);
  fa::fa_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(o_acc, lse_acc, o_new, lse_new, rows, head_dim / 4);
  FA_CUDA(cudaGetLastError());
  fa::fa_merge_lse_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(lse_acc, lse_new, rows);
  FA_CUDA(cudaGetLastError());
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return FA_OK;
}

int fa_cast_f32(const float* src, void* dst, int64_t n, int32_t dtype, void* stream) {
  if (!src || !dst || n <= 0 || (dtype != FA_BF16 && dtype != FA_F16)) return FA_ERR_INVALID_ARG;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  const int64_t threads = (n + 3) / 4;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == FA_BF16) fa::fa_cast_16_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  else fa::fa_cast_16_kernel<__half><<<grid, 256, 0, st>>>(src, static_cast<__half*>(dst), n);
  FA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}
int fa_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) { return fa_cast_f32(src, dst, n, FA_BF16, stream); }

// ------------------------------ peer-to-peer staging (ring forward) ------------------------------
// One process per GPU: a rank exports its K/V shard buffer as a CUDA IPC handle, every other rank maps it and pulls the
// shard it needs next with a plain device-to-device cudaMemcpyAsync — a copy-engine transfer over NVLink that needs no SM,
// so it runs under the persistent attention kernel (which owns every SM for the whole step).
int fa_p2p_alloc(int64_t bytes, void** ptr, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (bytes <= 0 || !ptr || !handle) return FA_ERR_INVALID_ARG;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  void* p = nullptr;
  FA_CUDA(cudaMalloc(&p, (size_t)bytes));   // its own allocation: the IPC handle then maps exactly this buffer at offset 0
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  memcpy(handle, &h, 64);
  *ptr = p;
  return FA_OK;
}
int fa_p2p_open(const uint8_t handle[64], void** peer_ptr) {
  if (!handle || !peer_ptr) return FA_ERR_INVALID_ARG;
  int major = 0;
  int rc = probe_device(&major);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  FA_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FA_OK;
}
int fa_p2p_close(void* peer_ptr) {
  if (!peer_ptr) return FA_ERR_INVALID_ARG;
  FA_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return FA_OK;
}
int fa_p2p_free(void* ptr) {
  if (!ptr) return FA_ERR_INVALID_ARG;
  FA_CUDA(cudaFree(ptr));
  return FA_OK;
}
int fa_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  if (!dst || !src || bytes <= 0) return FA_ERR_INVALID_ARG;
  FA_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return FA_OK;
}

// ------------------------------ reference-named shims ------------------------------
static void die_on(int rc, const char* where) {
  if (rc != FA_OK) {
    printf("[fa_b200 ERROR] %s: %s %s\n", where, fa_strerror(rc), fa_last_cuda_error());
    exit(EXIT_FAILURE);
  }
}
static void sync_or_die(const char* where) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[CUDA ERROR] at %s:\n%s\n", where, cudaGetErrorString(e));
    exit(EXIT_FAILURE);
  }
}

void run_flash_tiled_coarse(float* O, float* K_d, float* Q_d, float* V_d, int batch_size, int seq_len) {
  die_on(fa_forward(Q_d, K_d, V_d, O, nullptr, 1, batch_size, seq_len, seq_len, 64, 1.0f, 0, FA_F32, nullptr), "run_flash_tiled_coarse");
  sync_or_die("run_flash_tiled_coarse");
}
void run_flash_tiled_coarse_causal(float* O, float* K_d, float* Q_d, float* V_d, int batch_size, int seq_len) {
  die_on(fa_forward(Q_d, K_d, V_d, O, nullptr, 1, batch_size, seq_len, seq_len, 64, 1.0f, 1, FA_F32, nullptr),
         "run_flash_tiled_coarse_causal");
  sync_or_die("run_flash_tiled_coarse_causal");
}
void attention_forward6(float* out, const float* inp, int B, int T, int C, int NH, const int block_size) {
  (void)block_size;  // only sized the reference's permute/unpermute launches, which no longer exist
  const int hs = C / NH;
  // the llm.c harness validates `out` against its CPU loop at 1e-4 (src/llm.c/attention_forward.cu:1262): fp32-grade
  // contractions by default; FA_B200_LLMC_TF32=1 selects the plain tf32 instance (faster, ~3e-4)
  static const bool fast = [] { const char* e = getenv("FA_B200_LLMC_TF32"); return e && atoi(e) != 0; }();
  die_on(fa_forward_packed_qkv_ex(inp, out, nullptr, B, T, NH, hs, 1.0f / sqrtf((float)hs), 1, fast ? 0 : FA_FLAG_PRECISE, nullptr),
         "attention_forward6");
  sync_or_die("attention_forward6");
}
void attention_forward(int kernel_num, float* out, float* vaccum, float* qkvr, float* preatt, float* att, const float* inp, int B,
                       int T, int C, int NH, const int block_size) {
  (void)vaccum; (void)qkvr; (void)preatt; (void)att;
  if (kernel_num != 6) {
    printf("Invalid kernel number\n");
    exit(1);
  }
  attention_forward6(out, inp, B, T, C, NH, block_size);
}

}  // extern "C"
