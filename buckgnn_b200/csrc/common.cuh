// Shared device helpers for the buckgnn_b200 kernels (sm_100a only).
//
// Thin inline-PTX wrappers for the Blackwell primitives the kernels use: mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / TMEM load / commit),
// cluster helpers, plus vector load/store and a watchdog for barrier waits so a
// protocol bug traps instead of hanging the GPU.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/buckgnn_b200.h"

#define BG_DEVINL __device__ __forceinline__

namespace bg {

constexpr int kHidden = 512;          // the only hidden width the tcgen05 path is built for
constexpr int kBigRowThreshold = 64;  // in-degree above which a row is a "hub" row

// ------------------------------------------------------------------ host-side helpers
#define BG_CUDA_OK(expr)                                  \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) { bg::set_last_cuda_error(_e, #expr, __FILE__, __LINE__); return BG_ERR_CUDA; } \
  } while (0)

#define BG_LAUNCH_OK()                                    \
  do {                                                    \
    cudaError_t _e = cudaGetLastError();                  \
    if (_e != cudaSuccess) { bg::set_last_cuda_error(_e, "kernel launch", __FILE__, __LINE__); return BG_ERR_CUDA; } \
  } while (0)

void set_last_cuda_error(cudaError_t e, const char* what, const char* file, int line);
int sm_count();

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ small device utils
BG_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
BG_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

BG_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

BG_DEVINL uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// ptxas reorders ordinary loads freely (under register pressure it SINKS a prefetch to its first use, exposing the
// whole DRAM latency), but it keeps .volatile operations of a thread in program order: a volatile load issued ahead of
// volatile shared-memory stores stays ahead of them.
BG_DEVINL uint4 ldg_volatile_v4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
// cached variant: gathers re-touch neighbouring rows, let L1 keep them
BG_DEVINL uint4 ldg_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
BG_DEVINL void stg_v4(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

BG_DEVINL float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
BG_DEVINL float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
BG_DEVINL uint32_t pack_bf16(float a, float b) {   // a -> low half, b -> high half, round-to-nearest-even
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

BG_DEVINL uint32_t pack_f16(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// sm_100 packed fp32 pairs (FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per instruction, each rounded exactly like
// the scalar instruction).  A pair lives in an aligned 64-bit register pair; pack / unpack are register renames.
BG_DEVINL uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
BG_DEVINL void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
BG_DEVINL uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
BG_DEVINL uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
BG_DEVINL uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

// 16-bit storage formats: unpack a pair / pack a pair / scalar convert / fp32 += 16-bit pair.
// add2 uses sm_100's mixed-precision add (add.rn.f32.{f16,bf16} -> one FHADD per element, the
// 16-bit half selected for free), so fp32 accumulation of 16-bit rows needs no conversions.
template <typename T> struct Pack16;
template <> struct Pack16<__nv_bfloat16> {
  static BG_DEVINL float lo(uint32_t u) { return bf16_lo(u); }
  static BG_DEVINL float hi(uint32_t u) { return bf16_hi(u); }
  static BG_DEVINL void add2(float& a0, float& a1, uint32_t u) {
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\t"
        "add.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}" : "+f"(a0), "+f"(a1) : "r"(u));
  }
  // out-of-place: fp32(low / high 16-bit half of u) + b  (b may come straight from a uniform register)
  static BG_DEVINL float add_lo(uint32_t u, float b) {
    float r; asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, lo, %2;\n\t}" : "=f"(r) : "r"(u), "f"(b)); return r;
  }
  static BG_DEVINL float add_hi(uint32_t u, float b) {
    float r; asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, hi, %2;\n\t}" : "=f"(r) : "r"(u), "f"(b)); return r;
  }
  static BG_DEVINL uint32_t pack(float a, float b) { return pack_bf16(a, b); }
  static BG_DEVINL __nv_bfloat16 one(float a) { return __float2bfloat16_rn(a); }
  static BG_DEVINL uint32_t hadd2(uint32_t a, uint32_t b) {          // packed pair + packed pair
    __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static BG_DEVINL uint32_t relu2(uint32_t a) {                        // max(pair, 0)
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), __floats2bfloat162_rn(0.f, 0.f));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static BG_DEVINL uint32_t hmax2(uint32_t a, uint32_t b) {            // packed max (exact: the result is one of the inputs)
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static constexpr uint32_t kNegInf2 = 0xff80ff80u;
};
template <> struct Pack16<__half> {
  static BG_DEVINL float lo(uint32_t u) { return __low2float(*reinterpret_cast<const __half2*>(&u)); }
  static BG_DEVINL float hi(uint32_t u) { return __high2float(*reinterpret_cast<const __half2*>(&u)); }
  static BG_DEVINL void add2(float& a0, float& a1, uint32_t u) {
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\t"
        "add.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;\n\t}" : "+f"(a0), "+f"(a1) : "r"(u));
  }
  static BG_DEVINL float add_lo(uint32_t u, float b) {
    float r; asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %2;\n\t}" : "=f"(r) : "r"(u), "f"(b)); return r;
  }
  static BG_DEVINL float add_hi(uint32_t u, float b) {
    float r; asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, hi, %2;\n\t}" : "=f"(r) : "r"(u), "f"(b)); return r;
  }
  static BG_DEVINL uint32_t pack(float a, float b) { return pack_f16(a, b); }
  static BG_DEVINL __half one(float a) { return __float2half_rn(a); }
  static BG_DEVINL uint32_t hadd2(uint32_t a, uint32_t b) {
    __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static BG_DEVINL uint32_t relu2(uint32_t a) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), __floats2half2_rn(0.f, 0.f));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static BG_DEVINL uint32_t hmax2(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static constexpr uint32_t kNegInf2 = 0xfc00fc00u;
};

// ------------------------------------------------------------------ watchdog
// Barrier waits are bounded: after ~2^31 cycles (about a second) the kernel records
// where it was stuck and traps, so a pipeline bug surfaces as a CUDA error, not a hang.
__device__ unsigned int g_watchdog_info[4];

BG_DEVINL void watchdog_trip(uint32_t tag, uint32_t a, uint32_t b) {
  g_watchdog_info[0] = tag; g_watchdog_info[1] = a; g_watchdog_info[2] = b; g_watchdog_info[3] = blockIdx.x;
  __threadfence_system();
  __trap();
}

// ------------------------------------------------------------------ mbarrier
BG_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
BG_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
BG_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

BG_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
BG_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of this cluster
BG_DEVINL void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 r;\n\t"
      "mapa.shared::cluster.u32 r, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [r];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}
BG_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
BG_DEVINL void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifdef BG_WAIT_SLEEP                     // experiment: back off between polls (idle-warp power); tag 2/3 = the MMA warp's waits
    if (tag != 2u && tag != 3u) __nanosleep(BG_WAIT_SLEEP);
#endif
    if ((++spins & 1023u) == 0 && clock64() - t0 > (1ll << 31)) watchdog_trip(tag, bar, parity);
  }
}

// ------------------------------------------------------------------ cluster
BG_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
BG_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ TMA
BG_DEVINL void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
// 2D tile load, completion on an mbarrier of this CTA
BG_DEVINL void tma_load_2d(uint32_t smem_dst, const void* desc, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// 2-CTA variant: data lands in this CTA's smem, bytes are counted on the LEADER CTA's barrier
BG_DEVINL void tma_load_2d_cg2(uint32_t smem_dst, const void* desc, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(desc), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
BG_DEVINL void tma_store_2d(const void* desc, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(desc), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
BG_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending> BG_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
template <int kPending> BG_DEVINL void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// ------------------------------------------------------------------ tcgen05
template <int kCg> BG_DEVINL void tmem_alloc(uint32_t smem_result, uint32_t ncols) {
  if constexpr (kCg == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCg> BG_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCg == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
BG_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
BG_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers bf16/fp16 inputs, kind::tf32 fp32-as-tf32
template <int kCg, bool kTf32>
BG_DEVINL void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCg == 1 && !kTf32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  else if constexpr (kCg == 2 && !kTf32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  else if constexpr (kCg == 1 && kTf32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// "all MMAs issued so far by this thread are done" -> arrive on an mbarrier.
// cg2: multicast to the barrier at the same offset in both CTAs of the pair.
template <int kCg> BG_DEVINL void umma_commit(uint32_t bar) {
  if constexpr (kCg == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i = TMEM lane base+i)
BG_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
BG_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart):
// start address >> 4 | SBO (1024 B) >> 4 at bit 32 | version 1 at bit 46 | SWIZZLE_128B (2) at bit 61.
BG_DEVINL uint64_t umma_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major, 128-byte-swizzled operand tile as TMA lays it down from a row-major [K rows, MN cols] matrix: boxes of
// {128 B of MN, kb rows of K}; a K row is 128 B, 8-row K groups are 1024 B apart (SBO), the next 128 B of MN is the
// next box, `lbo_bytes` = kb * 128 further (LBO).  (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units, mma_traits_sm100.hpp.)
// 32-bit (tf32) operands use the 32-byte-atom variant of the swizzle (TMA SWIZZLE_128B_ATOM_32B, UMMA layout type
// SWIZZLE_128B_BASE32B = 1): K groups of 4 rows, 512 B apart.
BG_DEVINL uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, bool base32) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((base32 ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base32 ? 1 : 2) << 61;
  return d;
}
// instruction descriptor: fp32 accumulate, A/B both K-major; operand format 0 = f16, 1 = bf16, 2 = tf32
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t m, uint32_t n) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <typename T> struct is_bf16 { static constexpr bool value = false; };
template <> struct is_bf16<__nv_bfloat16> { static constexpr bool value = true; };

}  // namespace bg
