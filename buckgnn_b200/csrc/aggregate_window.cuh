// K2, window variant (EXPERIMENT, off by default -- see the measured result at the end of this comment): neighbourhood
// aggregation with the band's rows staged ONCE through shared memory.
//
// Replaces the same reference code as aggregate.cuh (`x[src]` gather + `scatter_add_` + divide inside PyG
// SAGEConv.propagate, call sites Models/BuckGNN.py:342,393,434,449,463) for 16-bit 512-column rows.
//
// Why: in k_aggregate_rows every neighbour row is fetched through L1/L2 by the warp that needs it.  On the plate meshes
// the neighbours of row i are i+-1, i+-nx (and i+-nx+-1 on the stiffened plates): the same 1 KB row is requested ~5-9
// times, the requests that miss L1 wait on L2/HBM latency (ncu r01: 57 % long-scoreboard stalls, 4.7 TB/s at degree 5,
// 2.6 TB/s at degree 11), and the achieved HBM rate depends on how well the L1 happens to hold the window.
// Here one producer warp streams the band's rows, in order and exactly once, by 1-D TMA bulk copies into a ring of
// kWinSlots x 8 rows (216 KB: the whole shared memory of the SM); a neighbour within kWinReach rows of the target row is
// read from the ring (LDS.128, no L1/L2 traffic, no miss latency), anything else (the super node's row, virtual edges,
// meshes wider than the reach) takes the global path of aggregate.cuh.  HBM sees one sequential read of x and one
// write of the aggregate -- the algorithmic minimum -- and the gather warps wait on shared memory, not on DRAM.
// The summation order is the CSR order in both paths, so the result is bit-identical to k_aggregate_rows.
//
// Range hubs (the reference's super node, see aggregate.cuh) are folded in as before, but the two hub warps now read
// the rows from the ring instead of streaming them a second time.
//
// Ring protocol: chunk c holds rows [base + 8c, base + 8c + 8), base = band start - kWinBack, slot c % kWinSlots, one
// mbarrier per slot (completed by the copy's transaction bytes).  Every consumer warp publishes the lowest row it still
// needs; the producer re-uses a slot only when that minimum has passed the slot's rows.  A warp working on row i waits
// for the chunk holding row i + kWinReach, so the fastest warp can be at most ~40 rows ahead of the slowest one.
//
// MEASURED (r02, B200, tools/agg_bench.py and tools/bench_configs.py cfg5; tests/test_gpu_kernels.py pass with
// BG_AGG_WINDOW=1, results bit-identical): cfg 2 (degree 5) 0.593 ms vs 0.422 ms for k_aggregate_rows; stiffened meshes
// (degree 11) 2.44 vs 2.07 ms per forward.  The window a row needs (+-84 rows = 21 chunks) leaves 6 of the 27 ring slots
// for copies in flight: 48 KB per SM against the ~66 KB that 44 GB/s per SM needs at ~1.5 us HBM latency, and the gather
// front advances in lock step with the stream.  A ring deep enough (window + 64 rows in flight + warp spread ~ 270 KB)
// does not fit one SM; k_aggregate_rows's L1 window + TMA-fed hub stream stays the product path.
#pragma once
#include "aggregate.cuh"

namespace bg {

constexpr int kWinChunk = 8;                         // rows per bulk copy (8 KB)
constexpr int kWinSlots = 27;                        // ring slots -> 216 rows resident
constexpr int kWinBack = 88;                         // rows staged before the band's first row (multiple of kWinChunk)
constexpr int kWinReach = 84;                        // |j - i| <= reach: neighbour j of row i is read from the ring
constexpr int kWinThreads = 768;                     // warp 0 producer, [1, 2 hub warps,] the rest gather warps
constexpr int kWinRowBytes = kHidden * 2;
constexpr int kWinSmemBytes = 1024 + kWinSlots * kWinChunk * kWinRowBytes;
static_assert(kWinReach <= kWinBack && kWinBack % kWinChunk == 0, "window geometry");
static_assert(kWinSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory a CTA may have");

template <typename T>
BG_DEVINL void win_frag_load(RowFrag<T>& f, const T* __restrict__ x, const unsigned char* ring, int64_t base, int64_t i,
                             int32_t j, int lane) {
  const int64_t d = (int64_t)j - i;
  if (d >= -kWinReach && d <= kWinReach) {           // warp-uniform: every lane looks at the same neighbour
    const int32_t rel = (int32_t)((int64_t)j - base);
    const int32_t slot = (rel >> 3) % kWinSlots;
    const uint4* p = reinterpret_cast<const uint4*>(ring + (size_t)slot * (kWinChunk * kWinRowBytes) + (size_t)(rel & 7) * kWinRowBytes);
    f.q[0] = p[lane];
    f.q[1] = p[32 + lane];
  } else {
    f.load(x + (size_t)j * kHidden, lane);
  }
}

template <typename T, int kAggr, bool kFold>
__global__ void __launch_bounds__(kWinThreads, 1)
k_aggregate_window(const T* __restrict__ x, T* __restrict__ out, int64_t N, int64_t band,
                   const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const HubFold hf) {
  static_assert(sizeof(T) == 2, "16-bit rows only (a 216-row window of fp32 rows does not fit shared memory)");
  constexpr int kHubWarps = kFold ? StreamGeom<T>::kWarps : 0;
  // k_hub_finalize assumes chunk q of a band belongs to hub warp q % kWarps with StreamGeom's chunk size
  static_assert(StreamGeom<T>::kWarps == 1 || StreamGeom<T>::kChunk == kWinChunk, "window / finalize chunk ownership");
  constexpr int kConsumers = kWinThreads / 32 - 1;               // every warp but the producer
  constexpr int kGather = kConsumers - kHubWarps;
  extern __shared__ __align__(1024) unsigned char win_smem[];
  const uint32_t bars = smem_u32(win_smem);                       // [kWinSlots] mbarriers
  volatile int32_t* need = reinterpret_cast<volatile int32_t*>(win_smem + 512);   // [kConsumers]: lowest row still needed, relative to base
  unsigned char* ring = win_smem + 1024;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r_beg = (int64_t)blockIdx.x * band;
  const int64_t r_end = min(N, r_beg + band);
  if (r_beg >= r_end) return;
  const int64_t base = r_beg - kWinBack;
  const int32_t n_chunks = (int32_t)((r_end + kWinReach - base + kWinChunk - 1) / kWinChunk);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWinSlots; ++s) mbar_init(bars + 8u * s, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < kConsumers) need[threadIdx.x] = 0;
  __syncthreads();

  if (warp == 0) {
    // ================================================================ producer
    for (int32_t c = 0; c < n_chunks; ++c) {
      const int slot = c % kWinSlots;
      if (c >= kWinSlots) {                                       // the slot's previous rows: [.., (c - kWinSlots) * 8 + 7]
        const int32_t last_old = (c - kWinSlots) * kWinChunk + kWinChunk - 1;
        const long long t0 = clock64();
        for (uint32_t spins = 0;; ++spins) {
          int32_t v = (lane < kConsumers) ? need[lane] : 0x7fffffff;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
          if (v > last_old) break;
          __nanosleep(64);
          if ((spins & 1023u) == 1023u && clock64() - t0 > (1ll << 31)) watchdog_trip(79u, (uint32_t)c, (uint32_t)v);
        }
      }
      if (lane == 0) {
        const int64_t row0 = base + (int64_t)c * kWinChunk;
        const int64_t lo = max(row0, (int64_t)0), hi = min(row0 + kWinChunk, N);
        if (hi <= lo) {
          mbar_arrive(bars + 8u * slot);                          // nothing to copy: complete the phase for the waiters
        } else {
          const uint32_t bytes = (uint32_t)(hi - lo) * kWinRowBytes;
          mbar_arrive_expect_tx(bars + 8u * slot, bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(ring + (size_t)slot * (kWinChunk * kWinRowBytes) + (size_t)(lo - row0) * kWinRowBytes)),
                         "l"(x + (size_t)lo * kHidden), "r"(bytes), "r"(bars + 8u * slot) : "memory");
        }
      }
      __syncwarp();
    }
    return;
  }

  const int cw = warp - 1;                                        // consumer index
  int32_t c_ready = 0;                                            // chunks [0, c_ready) are known to have landed
  auto wait_chunks = [&](int32_t upto) {                          // inclusive
    upto = min(upto, n_chunks - 1);
    while (c_ready <= upto) {
      mbar_wait(bars + 8u * (uint32_t)(c_ready % kWinSlots), (uint32_t)(c_ready / kWinSlots) & 1u, 78u);
      ++c_ready;
    }
  };

  // Before a consumer warp leaves, every bulk copy must have landed (a copy in flight when the CTA exits would write into
  // shared memory that may already belong to another CTA).  The warp first drops its claim on the ring, so the producer can
  // issue the remaining chunks, then waits for the LAST phase of every slot (earlier phases are complete by then; waiting on
  // a phase that lies more than one behind would never return).
  auto drain = [&]() {
    if (lane == 0) need[cw] = 0x7fffffff;
    c_ready = max(c_ready, n_chunks - kWinSlots);
    wait_chunks(n_chunks - 1);
  };

  if constexpr (kFold) {
    if (cw < kHubWarps) {
      // ============================================================== hub warps: the band's own rows, from the ring
      // chunk cb (rows r_beg + 8 cb ..) belongs to hub warp cb % kHubWarps: the same ownership k_hub_finalize assumes
      const int sw = cw;
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      int32_t cur_h = -1;
      auto flush = [&]() {
        if (cur_h >= 0) {
          const int64_t slot = (int64_t)cur_h * hf.parts + ((int64_t)blockIdx.x - hf.hub_lo[cur_h] / band);
          float4* dst = reinterpret_cast<float4*>(hf.partial) + (slot * kHubWarps + sw) * 128 + lane;
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j * 32] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      };
      const int64_t n_rows = r_end - r_beg;
      const int64_t n_band_chunks = (n_rows + kWinChunk - 1) / kWinChunk;
      auto hub_ids = [&](int64_t cb) -> int32_t {
        const int64_t row = r_beg + cb * kWinChunk + lane;
        return (cb < n_band_chunks && lane < kWinChunk && row < r_end) ? hf.hub_of_row[row] : -1;
      };
      int32_t h_cur = hub_ids(sw), h_nxt = hub_ids(sw + kHubWarps);
      for (int64_t cb = sw; cb < n_band_chunks; cb += kHubWarps) {
        const int32_t c = (int32_t)cb + kWinBack / kWinChunk;       // ring chunk of band chunk cb
        if (lane == 0) need[cw] = c * kWinChunk;
        const int rows = (int)min((int64_t)kWinChunk, r_end - (r_beg + cb * kWinChunk));
        const int32_t h_nxt2 = hub_ids(cb + 2 * (int64_t)kHubWarps);
        const int32_t hmine = h_cur;
        h_cur = h_nxt; h_nxt = h_nxt2;
        wait_chunks(c);
        const uint4* buf = reinterpret_cast<const uint4*>(ring + (size_t)(c % kWinSlots) * (kWinChunk * kWinRowBytes));
        const bool uniform = rows == kWinChunk && __all_sync(0xffffffffu, lane >= kWinChunk || hmine == cur_h) && cur_h >= 0;
        if (uniform) {
#pragma unroll
          for (int k0 = 0; k0 < kWinChunk; k0 += 4) {
            RowFrag<T> f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { f[k].q[0] = buf[(k0 + k) * 64 + lane]; f[k].q[1] = buf[(k0 + k) * 64 + 32 + lane]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k].template accumulate<BG_AGGR_SUM>(acc);
          }
        } else {
          for (int k = 0; k < rows; ++k) {
            const int32_t h = __shfl_sync(0xffffffffu, hmine, k);
            if (h != cur_h) { flush(); cur_h = h; }
            if (h >= 0) {
              RowFrag<T> f;
              f.q[0] = buf[k * 64 + lane]; f.q[1] = buf[k * 64 + 32 + lane];
              f.template accumulate<BG_AGGR_SUM>(acc);
            }
          }
        }
        __syncwarp();
      }
      flush();
      drain();
      return;
    }
  }

  // ================================================================== gather warps
  const int gw = cw - kHubWarps;
  int64_t r = r_beg + gw;
  int32_t beg = 0, end = 0, my = 0, nbeg = 0, nend = 0, nmy = 0, n2beg = 0, n2end = 0;
  if (r < r_end) {
    beg = rowptr[r]; end = rowptr[r + 1];
    my = (lane < end - beg) ? col[beg + lane] : 0;
  }
  if (r + kGather < r_end) {
    nbeg = rowptr[r + kGather]; nend = rowptr[r + kGather + 1];
    nmy = (lane < nend - nbeg) ? col[nbeg + lane] : 0;
  }
  if (r + 2 * kGather < r_end) { n2beg = rowptr[r + 2 * kGather]; n2end = rowptr[r + 2 * kGather + 1]; }
  for (; r < r_end; r += kGather) {
    if (lane == 0) need[cw] = (int32_t)(r - kWinReach - base);      // rows below r - reach are no longer needed by this warp
    const int32_t n2my = (lane < n2end - n2beg) ? col[n2beg + lane] : 0;
    int32_t n3beg = 0, n3end = 0;
    if (r + 3 * kGather < r_end) { n3beg = rowptr[r + 3 * kGather]; n3end = rowptr[r + 3 * kGather + 1]; }
    const int32_t deg = end - beg;
    if (deg <= kBigRowThreshold) {                                  // hub rows: hub warps + k_hub_finalize, or k_aggregate_hubs
      wait_chunks((int32_t)((r + kWinReach - base) / kWinChunk));
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = agg_init<kAggr>();
      for (int32_t b0 = 0; b0 < deg; b0 += 32) {
        const int32_t cnt = min(32, deg - b0);
        const int32_t idx = (b0 == 0) ? my : ((lane < cnt) ? col[beg + b0 + lane] : 0);
        int32_t j = 0;
        for (; j + 4 <= cnt; j += 4) {
          RowFrag<T> f0, f1, f2, f3;
          win_frag_load(f0, x, ring, base, r, __shfl_sync(0xffffffffu, idx, j), lane);
          win_frag_load(f1, x, ring, base, r, __shfl_sync(0xffffffffu, idx, j + 1), lane);
          win_frag_load(f2, x, ring, base, r, __shfl_sync(0xffffffffu, idx, j + 2), lane);
          win_frag_load(f3, x, ring, base, r, __shfl_sync(0xffffffffu, idx, j + 3), lane);
          f0.template accumulate<kAggr>(acc);
          f1.template accumulate<kAggr>(acc);
          f2.template accumulate<kAggr>(acc);
          f3.template accumulate<kAggr>(acc);
        }
        for (; j < cnt; ++j) {
          RowFrag<T> f;
          win_frag_load(f, x, ring, base, r, __shfl_sync(0xffffffffu, idx, j), lane);
          f.template accumulate<kAggr>(acc);
        }
      }
      agg_finalize<kAggr, false>(acc, deg);
      RowFrag<T>::store(out + (size_t)r * kHidden, lane, acc);
    }
    beg = nbeg; end = nend; my = nmy; nbeg = n2beg; nend = n2end; nmy = n2my; n2beg = n3beg; n2end = n3end;
  }
  drain();
}

}  // namespace bg
