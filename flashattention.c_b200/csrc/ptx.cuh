// ptx.cuh — thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st / fences) and the shared-memory matrix descriptors.
// Nothing here is library code: every wrapper is one PTX instruction (or a spin loop around one).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fa {

#define FA_DEVINL __device__ __forceinline__

// ----------------------------------------------------------------------------------------------
// watchdog: every mbarrier wait is bounded.  On expiry the CTA records where it was stuck and traps,
// which surfaces as a CUDA error on the host instead of a hung GPU.
// ----------------------------------------------------------------------------------------------
__device__ unsigned int g_fa_watchdog[4];  // [0]=tag, [1]=blockIdx.x, [2]=threadIdx.x, [3]=parity
// optional copy in host-mapped pinned memory (set by the host side): still readable after the trap has killed the context
__device__ unsigned int* g_fa_watchdog_host = nullptr;

#ifndef FA_WATCHDOG_SPINS
#define FA_WATCHDOG_SPINS (1u << 22)
#endif

FA_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---------------------------------------- mbarrier --------------------------------------------
FA_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
FA_DEVINL void fence_mbar_init() {
  // make the inits visible to the async proxy (TMA, tcgen05.commit)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
FA_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
FA_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
// FA_WAIT_HINT_NS > 0: pass a suspend-time hint, so a waiting warp sleeps in hardware instead of re-issuing the
// try_wait (A/B aid: does a spinning warp take issue slots from the softmax warp on the same sub-partition?)
#ifndef FA_WAIT_HINT_NS
#define FA_WAIT_HINT_NS 0
#endif
FA_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if FA_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(static_cast<uint32_t>(FA_WAIT_HINT_NS))
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// Cold path of mbar_wait, kept out of line: every wait site is then a handful of instructions, which matters for the
// single-warp roles (MMA issuer, producer) whose code is evicted from the 6 KB L0 instruction cache by the softmax warps
// between two of their loop iterations.
#ifndef FA_WATCHDOG_INLINE
#define FA_WATCHDOG_INLINE 0   // 1: the cold path inlined at every wait site (A/B aid)
#endif
#if FA_WATCHDOG_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void mbar_watchdog_expired(uint32_t tag, uint32_t parity) {
  g_fa_watchdog[0] = tag;
  g_fa_watchdog[1] = blockIdx.x;
  g_fa_watchdog[2] = threadIdx.x;
  g_fa_watchdog[3] = parity;
  unsigned int* wh = g_fa_watchdog_host;
  if (wh != nullptr && atomicCAS(wh, 0u, tag) == 0u) {   // first expiry wins
    wh[1] = blockIdx.x;
    wh[2] = threadIdx.x;
    wh[3] = parity;
  }
  __threadfence_system();
  __trap();
}
FA_DEVINL void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag) {
  if (mbar_try_wait(bar, parity)) return;  // fast path: already complete
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > FA_WATCHDOG_SPINS) mbar_watchdog_expired(tag, parity);
  }
}

// ------------------------------------------ TMA -----------------------------------------------
FA_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load: coordinates are (c0 = innermost element, c1 = row, c2 = head, c3 = batch)
FA_DEVINL void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
FA_DEVINL void tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
FA_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
FA_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
FA_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to smem -> visible to the async proxy (TMA store)
FA_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------- tcgen05 ---------------------------------------------
FA_DEVINL void tmem_alloc(uint32_t smem_result, uint32_t ncols) {  // whole warp, ncols power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols)
               : "memory");
}
FA_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
FA_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
FA_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FA_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
FA_DEVINL void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
FA_DEVINL void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrives (count 1) on `bar` once every tcgen05.mma issued so far by this thread has completed;
// implies tcgen05.fence::before_thread_sync.
FA_DEVINL void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]
template <bool kTF32>
FA_DEVINL void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem desc]
template <bool kTF32>
FA_DEVINL void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// Instruction descriptor (upper 32 bits of the PTX idesc operand; bit layout of the UMMA instruction
// descriptor): [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 = f16, 1 = bf16, 2 = tf32)
// [15] A major (0 = K)  [16] B major (0 = K, 1 = MN)  [17,23) N >> 3  [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_format, uint32_t b_mn_major, uint32_t M, uint32_t N) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | (0u << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// Shared-memory matrix descriptor, SWIZZLE_128B layouts written by TMA (row pitch 128 B, 8-row
// groups of 1024 B):  [0,14) start address >> 4   [16,30) leading-dim byte offset >> 4
// [32,46) stride-dim byte offset >> 4   [46,48) version = 1 (sm_100)   [61,64) layout type.
__host__ __device__ constexpr uint64_t make_sdesc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | (static_cast<uint64_t>(layout_type) << 61);
}
constexpr uint32_t kLayoutSw128 = 2;         // SWIZZLE_128B: 16-byte units XOR-swizzled over 8 rows (TMA SWIZZLE_128B)
constexpr uint32_t kLayoutSw128Base32 = 1;   // SWIZZLE_128B_BASE32B: 32-byte units over 4 rows (TMA SWIZZLE_128B_ATOM_32B);
                                             // the only layout tcgen05 accepts for MN-major 32-bit (tf32) operands
__host__ __device__ constexpr uint64_t make_sdesc_hi_sw128(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_sdesc_hi(lbo_bytes, sbo_bytes, kLayoutSw128);
}
FA_DEVINL uint64_t sdesc_at(uint64_t hi_template, uint32_t smem_addr) {
  return hi_template | static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
}

// ---- TMEM <-> registers, 32 lanes x 32 bit pattern: thread t of warp w touches lane 32*(w%4)+t ----
FA_DEVINL void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
FA_DEVINL void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
FA_DEVINL void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
FA_DEVINL void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ------------------------------------------ misc ----------------------------------------------
// exactly one lane of a fully converged warp gets `true` (the same lane every time)
FA_DEVINL bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
FA_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 — two fp32 lanes per instruction)
FA_DEVINL float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
FA_DEVINL float2 fadd2(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
FA_DEVINL float2 fmul2(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
// exp2 on the FMA/ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax
// polynomial (max relative error 7.5e-5, well under the bf16 / tf32 quantisation of P), 2^n by adding n to the
// exponent field.  The magic constant 1.5 * 2^23 leaves round(x) in the low mantissa bits of (x + magic).
constexpr float kExp2Magic = 12582912.0f;
constexpr float kExp2C0 = 9.999280572e-01f, kExp2C1 = 6.932609677e-01f, kExp2C2 = 2.426111251e-01f, kExp2C3 = 5.517166853e-02f;
FA_DEVINL float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float xf = x + kExp2Magic;
  const float n = xf - kExp2Magic;
  const float f = x - n;
  float p = fmaf(kExp2C3, f, kExp2C2);
  p = fmaf(p, f, kExp2C1);
  p = fmaf(p, f, kExp2C0);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
FA_DEVINL float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 xf = fadd2(x, make_float2(kExp2Magic, kExp2Magic));
  const float2 n = fadd2(xf, make_float2(-kExp2Magic, -kExp2Magic));
  const float2 f = ffma2(n, make_float2(-1.0f, -1.0f), x);
  float2 p = ffma2(make_float2(kExp2C3, kExp2C3), f, make_float2(kExp2C2, kExp2C2));
  p = ffma2(p, f, make_float2(kExp2C1, kExp2C1));
  p = ffma2(p, f, make_float2(kExp2C0, kExp2C0));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(xf.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(xf.y) << 23));
  return r;
}
FA_DEVINL uint32_t pack_bf16x2(float lo, float hi) {  // result: low 16 bits = bf16(lo), high = bf16(hi)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
FA_DEVINL uint32_t pack_f16x2(float lo, float hi) {   // low 16 bits = fp16(lo), high = fp16(hi)
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool kF16>
FA_DEVINL uint32_t pack_16x2(float lo, float hi) {     // the 16-bit operand type of the kind::f16 instances: fp16 or bf16
  if constexpr (kF16) return pack_f16x2(lo, hi);
  else return pack_bf16x2(lo, hi);
}
FA_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
FA_DEVINL void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
FA_DEVINL void st_shared_b32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
FA_DEVINL uint32_t ld_shared_b32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
FA_DEVINL void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
FA_DEVINL void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

}  // namespace fa
