// K2: neighbourhood aggregation over the CSR of K1 (mean / sum / max), 512 columns.
//
// Replaces `x.index_select(0, src)` + `scatter_add_` + `/ clamp(count, 1)` inside
// PyG SAGEConv.propagate (reference call sites Models/BuckGNN.py:342,393,434,449,463).
// HBM/L2-bandwidth bound: every source row is read once from HBM (re-reads by the
// ~6 rows that share it hit L2/L1), the aggregate row is written once.
//
//  * normal rows: one warp per row, each SM walking a contiguous band of rows so the
//    mesh's neighbour reuse is served by L1; a lane owns 16 of the 512 columns and reads
//    them with 128-bit loads, so one warp-wide load instruction covers 512
//    contiguous bytes of a source row; neighbour indices are fetched 32 at a time
//    and broadcast by shuffle; 4 neighbours are in flight per lane.  fp32
//    accumulation in CSR order = ascending edge id (deterministic).
//  * hub rows (degree > 64: the super node, degree = graph size): split into
//    kHubSlices slices, one CTA per (hub, slice) writes an fp32 partial; the last
//    CTA to finish a hub (atomic ticket) reduces the partials in slice order and
//    writes the row -- no warp ever walks a 4k-32k neighbour list alone, and the
//    result does not depend on scheduling.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace bg {

constexpr int kHubSlices = 16;
constexpr int kAggWarpsPerBlock = 8;       // hub kernel
// row kernel: one CTA per SM; 32 warps for 16-bit rows (64 regs/thread), 16 warps for fp32 rows
#ifndef BG_AGG_THREADS16
#define BG_AGG_THREADS16 768
#endif
template <typename T> __host__ __device__ constexpr int agg_row_threads() { return sizeof(T) == 2 ? BG_AGG_THREADS16 : 512; }
template <typename T> __host__ __device__ constexpr int agg_fold_threads() { return agg_row_threads<T>(); }


template <int kAggr> BG_DEVINL float agg_init() { return kAggr == BG_AGGR_MAX ? -INFINITY : 0.f; }
template <int kAggr> BG_DEVINL float agg_op(float a, float b) {
  if constexpr (kAggr == BG_AGGR_MAX) return fmaxf(a, b); else return a + b;
}

// ---- per-lane row fragments: 16 values of one 512-wide row -------------------------------
// bf16/f16: two 16-byte chunks at columns [8*lane, +8) and [256 + 8*lane, +8)
// f32 : four 16-byte chunks at columns [4*lane + 128*j, +4), j = 0..3
template <typename T> struct RowFrag;

template <typename T> struct RowFrag16 {
  uint4 q[2];
  BG_DEVINL void load(const T* row, int lane) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
    q[0] = ldg_v4(p + lane);
    q[1] = ldg_v4(p + 32 + lane);
  }
  template <int kAggr> BG_DEVINL void accumulate(float (&acc)[16]) const {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (kAggr == BG_AGGR_MAX) {
        acc[2 * i] = fmaxf(acc[2 * i], Pack16<T>::lo(u[i]));
        acc[2 * i + 1] = fmaxf(acc[2 * i + 1], Pack16<T>::hi(u[i]));
      } else {
        Pack16<T>::add2(acc[2 * i], acc[2 * i + 1], u[i]);      // fp32 += 16-bit, one FHADD each
      }
    }
  }
  static BG_DEVINL void store(T* row, int lane, const float (&v)[16]) {
    uint4 a, b;
    a.x = Pack16<T>::pack(v[0], v[1]);  a.y = Pack16<T>::pack(v[2], v[3]);
    a.z = Pack16<T>::pack(v[4], v[5]);  a.w = Pack16<T>::pack(v[6], v[7]);
    b.x = Pack16<T>::pack(v[8], v[9]);  b.y = Pack16<T>::pack(v[10], v[11]);
    b.z = Pack16<T>::pack(v[12], v[13]); b.w = Pack16<T>::pack(v[14], v[15]);
    uint4* p = reinterpret_cast<uint4*>(row);
    stg_v4(p + lane, a);
    stg_v4(p + 32 + lane, b);
  }
  // column of accumulator slot i for this lane
  static BG_DEVINL int col_of(int lane, int i) { return (i < 8) ? (8 * lane + i) : (256 + 8 * lane + (i - 8)); }
};
template <> struct RowFrag<__nv_bfloat16> : RowFrag16<__nv_bfloat16> {};
template <> struct RowFrag<__half> : RowFrag16<__half> {};

template <> struct RowFrag<float> {
  uint4 q[4];
  BG_DEVINL void load(const float* row, int lane) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = ldg_v4(p + 32 * j + lane);
  }
  template <int kAggr> BG_DEVINL void accumulate(float (&acc)[16]) const {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = agg_op<kAggr>(acc[i], __uint_as_float(u[i]));
  }
  static BG_DEVINL void store(float* row, int lane, const float (&v)[16]) {
    uint4* p = reinterpret_cast<uint4*>(row);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 a;
      a.x = __float_as_uint(v[4 * j]); a.y = __float_as_uint(v[4 * j + 1]);
      a.z = __float_as_uint(v[4 * j + 2]); a.w = __float_as_uint(v[4 * j + 3]);
      stg_v4(p + 32 * j + lane, a);
    }
  }
  static BG_DEVINL int col_of(int lane, int i) { return 128 * (i >> 2) + 4 * lane + (i & 3); }
};

// accumulate rows col[beg..end) of x into acc, in order
template <typename T, int kAggr>
BG_DEVINL void gather_range(const T* __restrict__ x, const int32_t* __restrict__ col, int32_t beg, int32_t end,
                            int lane, float (&acc)[16]) {
  for (int32_t base = beg; base < end; base += 32) {
    const int32_t cnt = min(32, end - base);
    const int32_t my = (lane < cnt) ? col[base + lane] : 0;
    int32_t j = 0;
    for (; j + 4 <= cnt; j += 4) {
      RowFrag<T> f0, f1, f2, f3;
      f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
      f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 1) * kHidden, lane);
      f2.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 2) * kHidden, lane);
      f3.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 3) * kHidden, lane);
      f0.template accumulate<kAggr>(acc);
      f1.template accumulate<kAggr>(acc);
      f2.template accumulate<kAggr>(acc);
      f3.template accumulate<kAggr>(acc);
    }
    for (; j < cnt; ++j) {
      RowFrag<T> f;
      f.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
      f.template accumulate<kAggr>(acc);
    }
  }
}

template <int kAggr, bool kExactDiv> BG_DEVINL void agg_finalize(float (&acc)[16], int32_t deg) {
  if constexpr (kAggr == BG_AGGR_MEAN) {
    const float d = (float)max(deg, 1);
    if constexpr (kExactDiv) {                            // fp32 rows: true division, as `sum / count` does
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = acc[i] / d;
    } else {                                              // 16-bit rows: the 1-ulp difference is below the output rounding
      const float rd = 1.f / d;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] *= rd;
    }
  } else if constexpr (kAggr == BG_AGGR_MAX) {
    if (deg == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    }
  }
}

// One persistent 32-warp CTA per SM owns a CONTIGUOUS band of rows and walks it in order,
// 32 rows at a time.  Mesh neighbours of row i are i+-1 and i+-nx, so the band's reuse
// window (~2*nx+32 rows of 1 KB) stays in the SM's L1: a source row is fetched from L2
// once and hit ~3 more times, instead of every gather going to L2.
// gather the rows whose indices sit in `my` (lane j holds neighbour j, cnt <= 32 of them), 4 rows in flight;
// the last 1-3 rows are issued together as one more group (one exposed latency, not one per row).
// (A generic kBatch-wide version with predicated loads into a fragment array was 2x slower: ptxas kept the
// array in local memory -- tools/agg_bench.py, profiles/r01_agg_variants.txt.)
template <typename T, int kAggr>
BG_DEVINL void gather_indexed(const T* __restrict__ x, int32_t my, int32_t cnt, int lane, float (&acc)[16]) {
  int32_t j = 0;
  for (; j + 4 <= cnt; j += 4) {
    RowFrag<T> f0, f1, f2, f3;
    f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
    f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 1) * kHidden, lane);
    f2.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 2) * kHidden, lane);
    f3.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 3) * kHidden, lane);
    f0.template accumulate<kAggr>(acc);
    f1.template accumulate<kAggr>(acc);
    f2.template accumulate<kAggr>(acc);
    f3.template accumulate<kAggr>(acc);
  }
  const int32_t rem = cnt - j;
  if (rem == 3) {
    RowFrag<T> f0, f1, f2;
    f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
    f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 1) * kHidden, lane);
    f2.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 2) * kHidden, lane);
    f0.template accumulate<kAggr>(acc);
    f1.template accumulate<kAggr>(acc);
    f2.template accumulate<kAggr>(acc);
  } else if (rem == 2) {
    RowFrag<T> f0, f1;
    f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
    f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 1) * kHidden, lane);
    f0.template accumulate<kAggr>(acc);
    f1.template accumulate<kAggr>(acc);
  } else if (rem == 1) {
    RowFrag<T> f0;
    f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
    f0.template accumulate<kAggr>(acc);
  }
}

// Max aggregation over 16-bit rows stays PACKED: the maximum of 16-bit values is one of them, so HMNMX2 on the raw pairs
// gives bit for bit what converting every value to fp32, FMNMX and rounding back gave -- at 8 instead of 32 instructions
// per neighbour row and lane, with 8 instead of 16 accumulator registers and no conversion at the store.
template <typename T> struct PackedMax {
  uint4 m[2];
  BG_DEVINL void init() {
    m[0] = m[1] = make_uint4(Pack16<T>::kNegInf2, Pack16<T>::kNegInf2, Pack16<T>::kNegInf2, Pack16<T>::kNegInf2);
  }
  BG_DEVINL void take(const RowFrag<T>& f) {
    using P = Pack16<T>;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      m[j].x = P::hmax2(m[j].x, f.q[j].x); m[j].y = P::hmax2(m[j].y, f.q[j].y);
      m[j].z = P::hmax2(m[j].z, f.q[j].z); m[j].w = P::hmax2(m[j].w, f.q[j].w);
    }
  }
  BG_DEVINL void store(T* row, int lane, int32_t deg) const {      // a row without neighbours aggregates to 0, not -inf
    uint4* p = reinterpret_cast<uint4*>(row);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    stg_v4(p + lane, deg == 0 ? z : m[0]);
    stg_v4(p + 32 + lane, deg == 0 ? z : m[1]);
  }
};

// neighbours col[beg..end) (any count), 4 rows in flight
template <typename T>
BG_DEVINL void gather_range_max16(const T* __restrict__ x, const int32_t* __restrict__ col, int32_t beg, int32_t end,
                                  int lane, PackedMax<T>& acc) {
  for (int32_t base = beg; base < end; base += 32) {
    const int32_t cnt = min(32, end - base);
    const int32_t my = (lane < cnt) ? col[base + lane] : 0;
    int32_t j = 0;
    for (; j + 4 <= cnt; j += 4) {
      RowFrag<T> f0, f1, f2, f3;
      f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
      f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 1) * kHidden, lane);
      f2.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 2) * kHidden, lane);
      f3.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 3) * kHidden, lane);
      acc.take(f0); acc.take(f1); acc.take(f2); acc.take(f3);
    }
    for (; j < cnt; ++j) {
      RowFrag<T> f;
      f.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
      acc.take(f);
    }
  }
}

// the first <= 32 neighbours, whose indices sit in `my` (same grouping as gather_indexed)
template <typename T>
BG_DEVINL void gather_indexed_max16(const T* __restrict__ x, int32_t my, int32_t cnt, int lane, PackedMax<T>& acc) {
  int32_t j = 0;
  for (; j + 4 <= cnt; j += 4) {
    RowFrag<T> f0, f1, f2, f3;
    f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
    f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 1) * kHidden, lane);
    f2.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 2) * kHidden, lane);
    f3.load(x + (size_t)__shfl_sync(0xffffffffu, my, j + 3) * kHidden, lane);
    acc.take(f0); acc.take(f1); acc.take(f2); acc.take(f3);
  }
  const int32_t rem = cnt - j;
  if (rem > 0) {                                   // the last 1-3 rows issued together (slots past the end re-read row j)
    RowFrag<T> f0, f1, f2;
    f0.load(x + (size_t)__shfl_sync(0xffffffffu, my, j) * kHidden, lane);
    f1.load(x + (size_t)__shfl_sync(0xffffffffu, my, rem > 1 ? j + 1 : j) * kHidden, lane);
    f2.load(x + (size_t)__shfl_sync(0xffffffffu, my, rem > 2 ? j + 2 : j) * kHidden, lane);
    acc.take(f0); acc.take(f1); acc.take(f2);      // max is idempotent: a repeated row changes nothing
  }
}

// One persistent CTA per SM owns a CONTIGUOUS band of rows and walks it in order, one warp per row.
// Mesh neighbours of row i are i+-1 and i+-nx, so the band's reuse window (~2*nx+32 rows of 1 KB)
// stays in the SM's L1: a source row is fetched from L2 once and hit ~3 more times.
// A warp's rows form a dependent chain rowptr -> col -> x per row; it is software-pipelined two
// rows deep (the next row's neighbour indices and the row-after-next's offsets are loaded while
// the current row's feature rows are in flight), so one memory latency per row is exposed, not three.
//
// kFold (mean / sum only): "range hubs" (bg_csr_build: a hub row whose neighbours are exactly the
// contiguous rows lo..lo+deg-1 -- the reference's super node) are folded into this pass.  kStream of
// the CTA's warps do not gather: they stream the band's own rows once, in order, 4 rows in flight, and
// add each row to a register partial of the hub whose range contains it; the partial is flushed to
// hub_partial[hub][band - first band of the hub][stream warp] when the rows move on to another hub.
// They run beside the gather warps over the same band, so their reads are served by (or fill) the
// L1/L2 lines the gathers need anyway, and k_hub_finalize adds the partials in a fixed order.  The
// separate hub kernel's second pass over all of x (1 GB of HBM reads per layer at cfg 2) disappears.
struct HubFold {
  const int32_t* hub_of_row;   // [N]  hub slot of the range containing row r, or -1
  const int32_t* hub_lo;       // [n_big] first row of the hub's range
  float* partial;              // [n_big][parts][kStream][512] f32, lane-major
  int32_t parts;               // band slots per hub
};
#ifndef BG_AGG_STREAM_LEAD
#define BG_AGG_STREAM_LEAD 96                // rows the hub stream may run ahead of the gather front
#endif
static_assert(BG_AGG_STREAM_LEAD >= 0, "a negative lead can starve the stream warps");
constexpr int kStreamLead = BG_AGG_STREAM_LEAD;
constexpr int kMaxStreamWarps = 2;           // bound used by the workspace size (bg_hubfold_workspace_bytes)
// Stream geometry per storage type (r02 sweep, tools/gpu_r4b.sh on cfg 2, ms per launch): what matters is the bytes per
// bulk copy and in flight, not the row count.  16-bit rows: ONE stream warp, 16-row (16 KB) copies, 3 in flight
// 0.409 ms (two warps x 8 rows x 3: 0.422; 4-row copies 0.50; one warp x 8 rows x 6: 0.54; three warps 0.45).
// fp32 rows: one warp, 4-row (8 KB) copies, 6 in flight 1.03 ms (two warps x 8 rows x 3: 1.13; 16-row copies x 3: 1.45).
// A smaller ring would let the L1 window grow by a carve-out step (64 -> 32 KB), but every ring below ~48 KB lost more
// on the stream than the L1 gained.  -DBG_AGG_STREAM_WARPS / _CHUNK / _RING override all types (experiments).
template <typename T> struct StreamGeom {
#if defined(BG_AGG_STREAM_WARPS) || defined(BG_AGG_STREAM_CHUNK) || defined(BG_AGG_STREAM_RING)
#ifndef BG_AGG_STREAM_WARPS
#define BG_AGG_STREAM_WARPS 2
#endif
#ifndef BG_AGG_STREAM_CHUNK
#define BG_AGG_STREAM_CHUNK 8
#endif
#ifndef BG_AGG_STREAM_RING
#define BG_AGG_STREAM_RING 3
#endif
  static constexpr int kWarps = BG_AGG_STREAM_WARPS, kChunk = BG_AGG_STREAM_CHUNK, kRing = BG_AGG_STREAM_RING;
#else
  static constexpr int kWarps = 1;
  static constexpr int kChunk = sizeof(T) == 2 ? 16 : 4;   // rows per bulk copy (multiple of 4)
  static constexpr int kRing = sizeof(T) == 2 ? 3 : 6;     // bulk copies in flight per stream warp
#endif
  static_assert(kChunk % 4 == 0 && kChunk <= 32 && kWarps <= kMaxStreamWarps, "stream geometry");
};
template <typename T> constexpr int agg_fold_smem() {
  return 1024 + StreamGeom<T>::kWarps * StreamGeom<T>::kRing * StreamGeom<T>::kChunk * kHidden * (int)sizeof(T);
}
#ifndef BG_AGG_PF
#define BG_AGG_PF 0                          // 0: no prefetch (default), 1: prefetch.global.L2, 2: prefetch.global.L1 -- both slower
#endif
template <int kLevel> BG_DEVINL void prefetch_line(const void* p) {
  if constexpr (kLevel == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  else asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <typename T, int kAggr, bool kFold, int kThreads>
__global__ void __launch_bounds__(kThreads, 1)
k_aggregate_rows(const T* __restrict__ x, T* __restrict__ out, int64_t N, int64_t band,
                 const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const HubFold hf) {
  constexpr int kStream = kFold ? StreamGeom<T>::kWarps : 0;
  constexpr int kStreamChunk = StreamGeom<T>::kChunk, kStreamRing = StreamGeom<T>::kRing;
  constexpr int kWarps = kThreads / 32 - kStream;            // gather warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r_beg = (int64_t)blockIdx.x * band;
  const int64_t r_end = min(N, r_beg + band);
  extern __shared__ __align__(128) unsigned char agg_smem[];
  volatile int32_t* progress = reinterpret_cast<volatile int32_t*>(agg_smem);        // gather front (row offset in band)
  if constexpr (kFold) {
    constexpr uint32_t kRowBytes = kHidden * (uint32_t)sizeof(T);
    constexpr uint32_t kBufBytes = kStreamChunk * kRowBytes;
    const uint32_t bars = smem_u32(agg_smem) + 64;                                   // [kStream][kStreamRing] mbarriers
    unsigned char* bufs = agg_smem + 1024;                                           // [kStream][kStreamRing][kBufBytes]
    if (threadIdx.x == 0) {
      *progress = 0;
      for (int i = 0; i < kStream * kStreamRing; ++i) mbar_init(bars + 8u * i, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (warp >= kWarps) {
      // ---------------------------------------------------------------- hub streaming warps
      // Chunks of kStreamChunk consecutive rows arrive by 1-D bulk copies (TMA) into a small ring: no registers
      // are tied up by loads in flight.  The stream is throttled to kStreamLead rows ahead of the gather front
      // (warp 0's row, published in shared memory): ~one mesh row ahead it is the stream that takes the HBM
      // misses of the rows entering the gather window, and the gathers find them in L2; further ahead the lines
      // are evicted again before they are used (measured: lead 96 -> 0.42 ms, 384 -> 0.52 ms, profiles/r01_agg_variants.txt).
      const int sw = warp - kWarps;
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      int32_t cur_h = -1;
      auto flush = [&]() {
        if (cur_h >= 0) {
          const int64_t slot = (int64_t)cur_h * hf.parts + ((int64_t)blockIdx.x - hf.hub_lo[cur_h] / band);
          float4* dst = reinterpret_cast<float4*>(hf.partial) + (slot * kStream + sw) * 128 + lane;
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j * 32] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      };
      const int64_t n_rows = r_end - r_beg;
      const int64_t n_chunks = (n_rows + kStreamChunk - 1) / kStreamChunk;
      const uint32_t my_bars = bars + 8u * (uint32_t)(sw * kStreamRing);
      unsigned char* my_bufs = bufs + (size_t)sw * kStreamRing * kBufBytes;
      auto issue = [&](int64_t c, int slot) {                  // lane 0: chunk c -> ring slot
        const int64_t row0 = r_beg + c * kStreamChunk;
        const uint32_t bytes = (uint32_t)min((int64_t)kStreamChunk, r_end - row0) * kRowBytes;
        while ((c * kStreamChunk) > (int64_t)*progress + kStreamLead) __nanosleep(256);
        mbar_arrive_expect_tx(my_bars + 8u * slot, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(my_bufs + (size_t)slot * kBufBytes)), "l"(x + (size_t)row0 * kHidden), "r"(bytes),
                       "r"(my_bars + 8u * slot) : "memory");
      };
      if (lane == 0)
        for (int k = 0; k < kStreamRing; ++k)
          if (sw + (int64_t)k * kStream < n_chunks) issue(sw + (int64_t)k * kStream, k);
      // hub ids of this warp's chunks, fetched two chunks ahead (each is a fresh 32-byte sector from HBM)
      auto hub_ids = [&](int64_t c) -> int32_t {
        const int64_t row = r_beg + c * kStreamChunk + lane;
        return (c < n_chunks && lane < kStreamChunk && row < r_end) ? hf.hub_of_row[row] : -1;
      };
      int32_t h_cur = hub_ids(sw), h_nxt = hub_ids(sw + kStream);
      int64_t it = 0;
      for (int64_t c = sw; c < n_chunks; c += kStream, ++it) {
        const int slot = (int)(it % kStreamRing);
        const int64_t row0 = r_beg + c * kStreamChunk;
        const int rows = (int)min((int64_t)kStreamChunk, r_end - row0);
        const int32_t h_nxt2 = hub_ids(c + 2 * (int64_t)kStream);
        const int32_t hmine = h_cur;
        h_cur = h_nxt; h_nxt = h_nxt2;
        mbar_wait(my_bars + 8u * slot, (uint32_t)(it / kStreamRing) & 1u, 77u);
        const uint4* buf = reinterpret_cast<const uint4*>(my_bufs + (size_t)slot * kBufBytes);
        constexpr int kQ = (int)(sizeof(RowFrag<T>) / sizeof(uint4));            // 16-byte pieces per lane per row
        const bool uniform = rows == kStreamChunk && __all_sync(0xffffffffu, lane >= kStreamChunk || hmine == cur_h) && cur_h >= 0;
        if (uniform) {                                  // whole chunk inside the current hub's range: 4 rows at a time
#pragma unroll
          for (int k0 = 0; k0 < kStreamChunk; k0 += 4) {
            RowFrag<T> f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int j = 0; j < kQ; ++j) f[k].q[j] = buf[(size_t)(k0 + k) * (kRowBytes / 16) + 32 * j + lane];
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k].template accumulate<BG_AGGR_SUM>(acc);
          }
        } else {
          for (int k = 0; k < rows; ++k) {
            const int32_t h = __shfl_sync(0xffffffffu, hmine, k);
            if (h != cur_h) { flush(); cur_h = h; }
            if (h >= 0) {
              RowFrag<T> f;
              const uint4* rp = buf + (size_t)k * (kRowBytes / 16);
#pragma unroll
              for (int j = 0; j < kQ; ++j) f.q[j] = rp[32 * j + lane];
              f.template accumulate<BG_AGGR_SUM>(acc);
            }
          }
        }
        __syncwarp();
        const int64_t nc = c + (int64_t)kStreamRing * kStream;
        if (lane == 0 && nc < n_chunks) issue(nc, slot);
      }
      flush();
      return;
    }
  }
  // Software pipeline over the warp's rows (stride kWarps), three deep: the row being gathered has its
  // neighbour indices in `my`; the next row's indices (`nmy`) arrived an iteration ago and are used NOW to
  // prefetch its neighbour rows into L2/L1 (one prefetch instruction covers 4 rows: lane l -> line l%8 of
  // neighbour l/8), so the row's own gathers find them on chip; the row after that has its offsets, and its
  // indices are requested now.  A gather is stalled on HBM latency ~60 % of the time without this
  // (ncu long_scoreboard, profiles/r01_v6_*): only the ~1 of 5 loads that misses L1 goes to HBM, so the loads
  // in flight cover too few HBM bytes.
  int64_t r = r_beg + warp;
  int32_t beg = 0, end = 0, my = 0, nbeg = 0, nend = 0, nmy = 0, n2beg = 0, n2end = 0;
  if (r < r_end) {
    beg = rowptr[r]; end = rowptr[r + 1];
    my = (lane < end - beg) ? col[beg + lane] : 0;
  }
  if (r + kWarps < r_end) {
    nbeg = rowptr[r + kWarps]; nend = rowptr[r + kWarps + 1];
    nmy = (lane < nend - nbeg) ? col[nbeg + lane] : 0;
  }
  if (r + 2 * kWarps < r_end) { n2beg = rowptr[r + 2 * kWarps]; n2end = rowptr[r + 2 * kWarps + 1]; }
  for (; r < r_end; r += kWarps) {
    if constexpr (kFold) {
      if (warp == 0 && lane == 0) *progress = (int32_t)(r - r_beg);
    }
#if BG_AGG_PF
    {
      const int32_t ndeg = min(nend - nbeg, 8);
      const int32_t nb0 = __shfl_sync(0xffffffffu, nmy, lane >> 3);
      const int32_t nb1 = __shfl_sync(0xffffffffu, nmy, 4 + (lane >> 3));
      if ((lane >> 3) < ndeg) prefetch_line<BG_AGG_PF>(reinterpret_cast<const char*>(x + (size_t)nb0 * kHidden) + (lane & 7) * (kHidden * (int)sizeof(T) / 8));
      if (4 + (lane >> 3) < ndeg) prefetch_line<BG_AGG_PF>(reinterpret_cast<const char*>(x + (size_t)nb1 * kHidden) + (lane & 7) * (kHidden * (int)sizeof(T) / 8));
    }
#endif
    // stage 2: indices of the row after next (its offsets arrived an iteration ago)
    const int32_t n2my = (lane < n2end - n2beg) ? col[n2beg + lane] : 0;
    // stage 1: offsets of the row after that
    int32_t n3beg = 0, n3end = 0;
    if (r + 3 * kWarps < r_end) { n3beg = rowptr[r + 3 * kWarps]; n3end = rowptr[r + 3 * kWarps + 1]; }
    const int32_t deg = end - beg;
    if (deg <= kBigRowThreshold) {                          // hub rows: k_aggregate_hubs / k_hub_finalize
      if constexpr (kAggr == BG_AGGR_MAX && sizeof(T) == 2) {
        PackedMax<T> pm;
        pm.init();
        gather_indexed_max16<T>(x, my, min(deg, 32), lane, pm);
        if (deg > 32) gather_range_max16<T>(x, col, beg + 32, end, lane, pm);
        pm.store(out + (size_t)r * kHidden, lane, deg);
      } else {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = agg_init<kAggr>();
        gather_indexed<T, kAggr>(x, my, min(deg, 32), lane, acc);
        if (deg > 32) gather_range<T, kAggr>(x, col, beg + 32, end, lane, acc);
        agg_finalize<kAggr, sizeof(T) == 4>(acc, deg);
        RowFrag<T>::store(out + (size_t)r * kHidden, lane, acc);
      }
    }
    beg = nbeg; end = nend; my = nmy; nbeg = n2beg; nend = n2end; nmy = n2my; n2beg = n3beg; n2end = n3end;
  }
  if constexpr (kFold) {
    if (warp == 0 && lane == 0) *progress = 0x7fffffff - 2 * kStreamLead;     // band done: release the stream warps
  }
}

// One CTA of 128 threads per range hub: thread (j, lane) owns accumulator slots 4j..4j+3 of `lane`, adds the
// per-(band, stream warp) partials of k_aggregate_rows<kFold> in (band, warp) order -- only the slots whose warp
// actually had a row inside the hub's range were written -- and writes the hub's aggregate row.
template <typename T, int kAggr>
__global__ void __launch_bounds__(128)
k_hub_finalize(T* __restrict__ out, int64_t N, int64_t band, const int32_t* __restrict__ rowptr,
               const int32_t* __restrict__ big_rows, const HubFold hf) {
  constexpr int kStream = StreamGeom<T>::kWarps, kStreamChunk = StreamGeom<T>::kChunk;
  const int32_t b = blockIdx.x;
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t r = big_rows[b];
  const int64_t deg = rowptr[r + 1] - rowptr[r];
  const int64_t lo = hf.hub_lo[b], hi = lo + deg;
  const int64_t c0 = lo / band, c1 = (hi - 1) / band;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t c = c0; c <= c1; ++c) {
    const int64_t a = max(lo, c * band), e = min(min(hi, (c + 1) * band), N);
    const float4* src = reinterpret_cast<const float4*>(hf.partial) + (((int64_t)b * hf.parts + (c - c0)) * kStream) * 128 + j * 32 + lane;
    // stream warp w of band c owns the 8-row chunks q with q % kStream == w; it wrote a partial for this hub iff
    // one of its chunks intersects [a, e)
    const int64_t qa = (a - c * band) / kStreamChunk, qe = (e - 1 - c * band) / kStreamChunk;
#pragma unroll
    for (int w = 0; w < kStream; ++w) {
      const int64_t q0 = qa + (((w - qa) % kStream) + kStream) % kStream;
      if (q0 <= qe) {
        const float4 v = __ldcg(src + (int64_t)w * 128);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  if constexpr (kAggr == BG_AGGR_MEAN) {
    const float d = (float)max(deg, (int64_t)1);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = v[k] / d;
  }
  T* orow = out + (size_t)r * kHidden;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = RowFrag<T>::col_of(lane, 4 * j + k);
    if constexpr (sizeof(T) == 2) orow[c] = Pack16<T>::one(v[k]); else orow[c] = v[k];
  }
}

// grid = n_big * kHubSlices CTAs of 8 warps; partial [n_big][kHubSlices][512] f32; ticket [n_big] (zeroed)
// kCompact: the aggregate of hub b goes to row b of `out` ([n_big, 512], the fused SAGE layer's side buffer) instead of
// row big_rows[b] of the [N, 512] aggregate matrix
template <typename T, int kAggr, bool kCompact = false>
__global__ void __launch_bounds__(kAggWarpsPerBlock * 32)
k_aggregate_hubs(const T* __restrict__ x, T* __restrict__ out,
                 const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                 const int32_t* __restrict__ big_rows, int32_t n_big,
                 float* __restrict__ partial, int32_t* __restrict__ ticket) {
  __shared__ float red[kAggWarpsPerBlock][kHidden];
  __shared__ int32_t is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t hub = blockIdx.x / kHubSlices, slice = blockIdx.x % kHubSlices;
  if (hub >= n_big) return;
  const int32_t r = big_rows[hub];
  const int32_t beg = rowptr[r], deg = rowptr[r + 1] - beg;
  // slice -> contiguous neighbour range; warp -> contiguous sub-range (order-preserving split)
  const int32_t s_beg = beg + (int32_t)((int64_t)deg * slice / kHubSlices);
  const int32_t s_end = beg + (int32_t)((int64_t)deg * (slice + 1) / kHubSlices);
  const int32_t s_len = s_end - s_beg;
  const int32_t w_beg = s_beg + (int32_t)((int64_t)s_len * warp / kAggWarpsPerBlock);
  const int32_t w_end = s_beg + (int32_t)((int64_t)s_len * (warp + 1) / kAggWarpsPerBlock);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = agg_init<kAggr>();
  gather_range<T, kAggr>(x, col, w_beg, w_end, lane, acc);
#pragma unroll
  for (int i = 0; i < 16; ++i) red[warp][RowFrag<T>::col_of(lane, i)] = acc[i];
  __syncthreads();
  float* my_partial = partial + ((size_t)hub * kHubSlices + slice) * kHidden;
  for (int c = threadIdx.x; c < kHidden; c += blockDim.x) {
    float v = red[0][c];
#pragma unroll
    for (int w = 1; w < kAggWarpsPerBlock; ++w) v = agg_op<kAggr>(v, red[w][c]);
    my_partial[c] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&ticket[hub], 1) == kHubSlices - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float* hp = partial + (size_t)hub * kHubSlices * kHidden;
  T* orow = out + (size_t)(kCompact ? hub : r) * kHidden;
  for (int c = threadIdx.x; c < kHidden; c += blockDim.x) {
    float v = __ldcg(hp + c);
    for (int s = 1; s < kHubSlices; ++s) v = agg_op<kAggr>(v, __ldcg(hp + (size_t)s * kHidden + c));
    if constexpr (kAggr == BG_AGGR_MEAN) v = v / (float)max(deg, 1);
    if constexpr (sizeof(T) == 2) orow[c] = Pack16<T>::one(v); else orow[c] = v;
  }
  if (threadIdx.x == 0) ticket[hub] = 0;    // leave the ticket ready for the next launch
}


// ------------------------------------------------------------------ 128-column rows
// The same aggregation over [N,128] rows (the node encoder's hidden layer): its last Linear is linear, so
// mean/sum aggregation commutes with it and layer 0 can aggregate 128 instead of 512 columns
// (engine.py: folded first layer).  A row is 256 B (16-bit: 16 lanes x 16 B, two rows per warp) or
// 512 B (fp32: one row per warp).
template <typename T> struct Narrow {
  static constexpr int kCols = 128;
  static constexpr int kLanes = kCols * (int)sizeof(T) / 16;   // lanes per row: 16 or 32
  static constexpr int kVals = 16 / (int)sizeof(T);            // values per lane: 8 or 4
  static constexpr int kRowsPerWarp = 32 / kLanes;
};

template <typename T, int kAggr>
BG_DEVINL void narrow_accumulate(float (&acc)[Narrow<T>::kVals], const uint4& q) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (kAggr == BG_AGGR_MAX) {
        acc[2 * i] = fmaxf(acc[2 * i], Pack16<T>::lo(u[i]));
        acc[2 * i + 1] = fmaxf(acc[2 * i + 1], Pack16<T>::hi(u[i]));
      } else {
        Pack16<T>::add2(acc[2 * i], acc[2 * i + 1], u[i]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = agg_op<kAggr>(acc[i], __uint_as_float(u[i]));
  }
}

// the 16 bytes that leave an accumulator unchanged: zeros for sum / mean, -inf for max
template <typename T, int kAggr> BG_DEVINL uint4 narrow_neutral() {
  if constexpr (kAggr != BG_AGGR_MAX) return make_uint4(0u, 0u, 0u, 0u);
  else if constexpr (sizeof(T) == 4) return make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u);
  else if constexpr (is_bf16<T>::value) return make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
  else return make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u);
}

template <typename T> BG_DEVINL uint4 narrow_pack(const float (&v)[Narrow<T>::kVals]) {
  uint4 o;
  if constexpr (sizeof(T) == 2) {
    o.x = Pack16<T>::pack(v[0], v[1]); o.y = Pack16<T>::pack(v[2], v[3]);
    o.z = Pack16<T>::pack(v[4], v[5]); o.w = Pack16<T>::pack(v[6], v[7]);
  } else {
    o.x = __float_as_uint(v[0]); o.y = __float_as_uint(v[1]); o.z = __float_as_uint(v[2]); o.w = __float_as_uint(v[3]);
  }
  return o;
}

// accumulate rows col[beg..end) of the [*,128] matrix x; executed by one sub-warp of Narrow<T>::kLanes lanes
template <typename T, int kAggr>
BG_DEVINL void narrow_gather(const T* __restrict__ x, const int32_t* __restrict__ col, int32_t beg, int32_t end,
                             int sl, uint32_t mask, float (&acc)[Narrow<T>::kVals]) {
  constexpr int kLanes = Narrow<T>::kLanes;
  const char* xb = reinterpret_cast<const char*>(x) + (size_t)sl * 16;
  constexpr size_t kRowBytes = 128 * sizeof(T);
  for (int32_t base = beg; base < end; base += kLanes) {
    const int32_t cnt = min(kLanes, end - base);
    const int32_t my = (sl < cnt) ? col[base + sl] : 0;
    int32_t j = 0;
    for (; j + 4 <= cnt; j += 4) {
      const uint4 q0 = ldg_v4(xb + (size_t)__shfl_sync(mask, my, j, kLanes) * kRowBytes);
      const uint4 q1 = ldg_v4(xb + (size_t)__shfl_sync(mask, my, j + 1, kLanes) * kRowBytes);
      const uint4 q2 = ldg_v4(xb + (size_t)__shfl_sync(mask, my, j + 2, kLanes) * kRowBytes);
      const uint4 q3 = ldg_v4(xb + (size_t)__shfl_sync(mask, my, j + 3, kLanes) * kRowBytes);
      narrow_accumulate<T, kAggr>(acc, q0); narrow_accumulate<T, kAggr>(acc, q1);
      narrow_accumulate<T, kAggr>(acc, q2); narrow_accumulate<T, kAggr>(acc, q3);
    }
    for (; j < cnt; ++j)
      narrow_accumulate<T, kAggr>(acc, ldg_v4(xb + (size_t)__shfl_sync(mask, my, j, kLanes) * kRowBytes));
  }
}

// Same band walk and two-deep index pipeline as k_aggregate_rows, one SUB-warp (16 or 32 lanes) per row.
// Control flow is WARP-UNIFORM: the two sub-warps of a 16-bit warp walk rows r0 and r0 + 1 in lock step (trip counts are
// the maximum over the two), so every shuffle uses the full-warp mask with width = kLanes.  (Round 1 gave each sub-warp
// its own mask and trip counts: a shuffle with a run-time mask compiles to a MATCH / WARPSYNC / ENDCOLLECTIVE sequence
// and the sub-warps executed one after the other -- ncu: 267 warp instructions per row, issue slots 74 % busy at 19 % of
// the DRAM peak.)  Up to 8 neighbours are in flight per row; slots past a row's degree are predicated off.
template <typename T, int kAggr>
__global__ void __launch_bounds__(1024, 1)
k_aggregate_rows128(const T* __restrict__ x, T* __restrict__ out, int64_t N,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col) {
  using NW = Narrow<T>;
  constexpr int kLanes = NW::kLanes;
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / kLanes, sl = lane % kLanes;
  const int64_t band = (N + gridDim.x - 1) / gridDim.x;
  const int64_t r_beg = (int64_t)blockIdx.x * band;
  const int64_t r_end = min(N, r_beg + band);
  constexpr int kStride = 32 * NW::kRowsPerWarp;               // rows per CTA iteration (32 warps)
  const char* xb = reinterpret_cast<const char*>(x) + (size_t)sl * 16;
  constexpr size_t kRowBytes = 128 * sizeof(T);
  auto offsets = [&](int64_t row, int32_t& b, int32_t& e) {
    b = 0; e = 0;
    if (row < r_end) { b = rowptr[row]; e = rowptr[row + 1]; }
  };
  int64_t r0 = r_beg + warp * NW::kRowsPerWarp;                // warp-uniform; this sub-warp's row is r0 + sub
  int32_t beg, end, nbeg, nend;
  offsets(r0 + sub, beg, end);
  int32_t my = (sl < end - beg) ? col[beg + sl] : 0;
  offsets(r0 + sub + kStride, nbeg, nend);
  for (; r0 < r_end; r0 += kStride) {
    const int64_t r = r0 + sub;
    const int32_t nmy = (sl < nend - nbeg) ? col[nbeg + sl] : 0;
    int32_t n2beg, n2end;
    offsets(r + 2 * kStride, n2beg, n2end);
    const int32_t deg = end - beg;
    const int32_t cnt = (deg <= kBigRowThreshold) ? min(deg, kLanes) : 0;     // hub rows: k_aggregate_hubs128
    int32_t wcnt = cnt;
    if constexpr (kLanes == 16) wcnt = max(wcnt, __shfl_xor_sync(kFull, cnt, 16));
    float acc[NW::kVals];
#pragma unroll
    for (int i = 0; i < NW::kVals; ++i) acc[i] = agg_init<kAggr>();
    for (int32_t base = 0; base < wcnt; base += 8) {           // uniform trip count
      // kN neighbour slots, all loads in flight before the first add; a slot past this row's degree loads nothing and
      // adds the neutral element (no per-lane branches).  Two unrolled bodies: 5 slots (plain quad meshes: 4 mesh
      // neighbours + the super node) and 8.
      auto round = [&](auto kN) {
        constexpr int n = decltype(kN)::value;
        uint4 q[n];
#pragma unroll
        for (int k = 0; k < n; ++k) {
          const int32_t nb = __shfl_sync(kFull, my, (base + k) & (kLanes - 1), kLanes);
          q[k] = narrow_neutral<T, kAggr>();
          if (base + k < cnt) q[k] = ldg_v4(xb + (size_t)nb * kRowBytes);
        }
#pragma unroll
        for (int k = 0; k < n; ++k) narrow_accumulate<T, kAggr>(acc, q[k]);
      };
      if (wcnt - base <= 5) round(std::integral_constant<int, 5>{});
      else round(std::integral_constant<int, 8>{});
    }
    if (deg <= kBigRowThreshold && r < r_end) {
      if (deg > kLanes) {                                      // rare: more neighbours than index lanes (sub-warp masks)
        const uint32_t mask = (kLanes == 32) ? kFull : (0xffffu << (16 * sub));
        narrow_gather<T, kAggr>(x, col, beg + kLanes, end, sl, mask, acc);
      }
      if constexpr (kAggr == BG_AGGR_MEAN) {
        const float d = (float)max(deg, 1);
        if constexpr (sizeof(T) == 4) {                        // fp32 rows: true division, as `sum / count` does
#pragma unroll
          for (int i = 0; i < NW::kVals; ++i) acc[i] = acc[i] / d;
        } else {                                               // 16-bit rows: the 1-ulp difference is below the output rounding
          const float rd = 1.f / d;
#pragma unroll
          for (int i = 0; i < NW::kVals; ++i) acc[i] *= rd;
        }
      } else if constexpr (kAggr == BG_AGGR_MAX) {
        if (deg == 0) {
#pragma unroll
          for (int i = 0; i < NW::kVals; ++i) acc[i] = 0.f;
        }
      }
      stg_v4(reinterpret_cast<char*>(out) + (size_t)r * 128 * sizeof(T) + (size_t)sl * 16, narrow_pack<T>(acc));
    }
    beg = nbeg; end = nend; my = nmy; nbeg = n2beg; nend = n2end;
  }
}

// hub rows of the 128-column aggregation: same slicing / ticket scheme as k_aggregate_hubs
template <typename T, int kAggr>
__global__ void __launch_bounds__(kAggWarpsPerBlock * 32)
k_aggregate_hubs128(const T* __restrict__ x, T* __restrict__ out,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const int32_t* __restrict__ big_rows, int32_t n_big,
                    float* __restrict__ partial, int32_t* __restrict__ ticket) {
  using NW = Narrow<T>;
  __shared__ float red[kAggWarpsPerBlock * NW::kRowsPerWarp][128];
  __shared__ int32_t is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / NW::kLanes, sl = lane % NW::kLanes;
  const uint32_t mask = (NW::kLanes == 32) ? 0xffffffffu : (0xffffu << (16 * sub));
  const int32_t hub = blockIdx.x / kHubSlices, slice = blockIdx.x % kHubSlices;
  if (hub >= n_big) return;
  const int32_t r = big_rows[hub];
  const int32_t beg = rowptr[r], deg = rowptr[r + 1] - beg;
  const int32_t s_beg = beg + (int32_t)((int64_t)deg * slice / kHubSlices);
  const int32_t s_end = beg + (int32_t)((int64_t)deg * (slice + 1) / kHubSlices);
  constexpr int kParts = kAggWarpsPerBlock * NW::kRowsPerWarp;      // sub-warps per CTA
  const int part = warp * NW::kRowsPerWarp + sub;
  const int32_t s_len = s_end - s_beg;
  const int32_t w_beg = s_beg + (int32_t)((int64_t)s_len * part / kParts);
  const int32_t w_end = s_beg + (int32_t)((int64_t)s_len * (part + 1) / kParts);
  float acc[NW::kVals];
#pragma unroll
  for (int i = 0; i < NW::kVals; ++i) acc[i] = agg_init<kAggr>();
  narrow_gather<T, kAggr>(x, col, w_beg, w_end, sl, mask, acc);
#pragma unroll
  for (int i = 0; i < NW::kVals; ++i) red[part][sl * NW::kVals + i] = acc[i];
  __syncthreads();
  float* my_partial = partial + ((size_t)hub * kHubSlices + slice) * 128;
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {
    float v = red[0][c];
#pragma unroll
    for (int w = 1; w < kParts; ++w) v = agg_op<kAggr>(v, red[w][c]);
    my_partial[c] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&ticket[hub], 1) == kHubSlices - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float* hp = partial + (size_t)hub * kHubSlices * 128;
  T* orow = out + (size_t)r * 128;
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {
    float v = __ldcg(hp + c);
    for (int s2 = 1; s2 < kHubSlices; ++s2) v = agg_op<kAggr>(v, __ldcg(hp + (size_t)s2 * 128 + c));
    if constexpr (kAggr == BG_AGGR_MEAN) v = v / (float)max(deg, 1);
    if constexpr (sizeof(T) == 2) orow[c] = Pack16<T>::one(v); else orow[c] = v;
  }
  if (threadIdx.x == 0) ticket[hub] = 0;
}

}  // namespace bg
