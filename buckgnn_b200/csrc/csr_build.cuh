// K1: on-device CSR build = stable counting sort of the edge list by key row,
// plus graph offsets from the PyG `batch` vector.
//
// Replaces the index handling PyG does implicitly inside SAGEConv.propagate /
// torch_scatter.scatter_mean (reference call sites Models/BuckGNN.py:449, 561) and
// global_mean_pool (:274).  Integer work, HBM bound, bit-exact vs
// argsort(key, stable) by construction:
//   1. histogram of key         (one 32-bit atomic per edge)
//   2. exclusive scan           (three small kernels, block sums in between)
//   3. unordered fill           (atomic cursor per row)
//   4. per-row sort of the edge ids -> stable order; rows <= 64 by one thread
//      (insertion sort on an L1-resident segment), hub rows by a CTA-wide
//      ascending-only bitonic network in shared memory.
// Steps 3+4 cost less than a full radix sort because rows are tiny (mesh degree
// ~5) apart from one hub row per graph.
#pragma once
#include "common.cuh"

namespace bg {

constexpr int kScanItemsPerBlock = 4096;   // 1024 threads x 4
constexpr int kSortSmemElems = 57344;      // hub rows up to this degree sort in shared memory (224 KB of the 227 KB a CTA
                                           // may have; the network skips partners beyond n, so n need not be a power of 2)

// An edge whose KEY endpoint lies outside [0, N) is dropped and flagged (info bit 0) by the histogram and the fill; the
// OTHER endpoint is only read where `col` is written (the sort kernels), which flag it there and store 0 -- the host
// raises on the flag before any kernel dereferences `col` (engine.PendingGraphIndex.finish).  Histogram and fill thus
// read 8 instead of 16 bytes per edge.
BG_DEVINL bool key_ok(int64_t k, int64_t N) { return k >= 0 && k < N; }
BG_DEVINL int32_t checked_col(int64_t o, int64_t N, int32_t* __restrict__ info) {
  if (o < 0 || o >= N) { atomicOr(&info[0], 1); return 0; }
  return (int32_t)o;
}

__global__ void k_csr_hist(const int64_t* __restrict__ key, int64_t E, int64_t N,
                           int32_t* __restrict__ deg, int32_t* __restrict__ info) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    int64_t k = key[e];
    if (!key_ok(k, N)) { bad = true; continue; }
    atomicAdd(&deg[k], 1);
  }
  if (bad) atomicOr(&info[0], 1);
}

// block-wide exclusive scan of 4 items per thread; returns block total via smem
BG_DEVINL int32_t block_exclusive_scan(int32_t thread_sum, int32_t* smem_warp /*[32]*/, int32_t& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t v = thread_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) smem_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int32_t w = (lane < (int)(blockDim.x >> 5)) ? smem_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    smem_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  block_total = smem_warp[(blockDim.x >> 5) - 1];
  int32_t warp_off = warp ? smem_warp[warp - 1] : 0;
  return warp_off + v - thread_sum;   // exclusive prefix of this thread's sum
}

__global__ void __launch_bounds__(1024) k_scan_block_sums(const int32_t* __restrict__ deg, int64_t N,
                                                          int32_t* __restrict__ block_sums) {
  __shared__ int32_t sw[32];
  int64_t base = (int64_t)blockIdx.x * kScanItemsPerBlock + threadIdx.x * 4;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) if (base + i < N) s += deg[base + i];
  int32_t total;
  block_exclusive_scan(s, sw, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place (any count, chunked)
__global__ void __launch_bounds__(1024) k_scan_of_sums(int32_t* __restrict__ block_sums, int32_t n_blocks) {
  __shared__ int32_t sw[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int32_t base = 0; base < n_blocks; base += 1024) {
    int32_t i = base + threadIdx.x;
    int32_t v = (i < n_blocks) ? block_sums[i] : 0;
    int32_t total;
    int32_t ex = block_exclusive_scan(v, sw, total);
    int32_t carry = carry_s;
    if (i < n_blocks) block_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
}

// rowptr = exclusive scan(deg); cursor = rowptr; list the hub rows
__global__ void __launch_bounds__(1024) k_scan_apply(const int32_t* __restrict__ deg, int64_t N,
                                                     const int32_t* __restrict__ block_offs,
                                                     int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                                     int32_t* __restrict__ big_rows, int32_t* __restrict__ info,
                                                     int32_t max_big) {
  __shared__ int32_t sw[32];
  int64_t base = (int64_t)blockIdx.x * kScanItemsPerBlock + threadIdx.x * 4;
  int32_t d[4], s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { d[i] = (base + i < N) ? deg[base + i] : 0; s += d[i]; }
  int32_t total;
  int32_t run = block_offs[blockIdx.x] + block_exclusive_scan(s, sw, total);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < N) {
      rowptr[base + i] = run;
      cursor[base + i] = run;
      if (d[i] > kBigRowThreshold) {
        int32_t slot = atomicAdd(&info[1], 1);
        if (slot < max_big) big_rows[slot] = (int32_t)(base + i);
        atomicMax(&info[5], d[i]);
      }
      if (d[i] == 0) atomicOr(&info[6], 1);               // some row has no entries (isolated node)
      run += d[i];
      if (base + i == N - 1) rowptr[N] = run;
    }
  }
}

__global__ void k_csr_fill(const int64_t* __restrict__ key, int64_t E, int64_t N,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    int64_t k = key[e];
    if (!key_ok(k, N)) continue;
    int32_t pos = atomicAdd(&cursor[k], 1);
    perm[pos] = (int32_t)e;
  }
}

// one thread per small row: insertion sort of its edge ids, then col = other[perm]
__global__ void k_csr_sort_small(const int32_t* __restrict__ rowptr, int64_t N,
                                 const int64_t* __restrict__ other, int32_t* __restrict__ perm,
                                 int32_t* __restrict__ col, int32_t* __restrict__ info) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int32_t s = rowptr[r], e = rowptr[r + 1];
  const int32_t d = e - s;
  if (d > kBigRowThreshold) return;
  for (int32_t i = s + 1; i < e; ++i) {
    int32_t v = perm[i];
    int32_t j = i - 1;
    while (j >= s && perm[j] > v) { perm[j + 1] = perm[j]; --j; }
    perm[j + 1] = v;
  }
  for (int32_t i = s; i < e; ++i) col[i] = checked_col(other[perm[i]], N, info);
}

// ascending-only bitonic network over n (<= P = pow2) keys; partners beyond n act as +inf
BG_DEVINL void bitonic_ascending(int32_t* a, int32_t n, int32_t P) {
  for (int32_t k = 2; k <= P; k <<= 1) {
    for (int32_t j = k >> 1; j > 0; j >>= 1) {
      for (int32_t i = threadIdx.x; i < P; i += blockDim.x) {
        int32_t l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < n) {
          int32_t x = a[i], y = a[l];
          if (x > y) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// A hub row whose sorted neighbour list is exactly lo, lo+1, ..., lo+n-1 (the super node: hub edges (i -> n)
// for i = 0..n-1 in order, reference VirtualEdgeCreate.py:106-111) is a "range hub": its aggregate is a sum
// over a contiguous band of rows, which the row kernel of the aggregation can accumulate as a by-product
// instead of re-reading every row (aggregate.cuh).  hub_lo[b] = lo or -1; hub_of_row[j] = b for j in the range;
// info[4] counts hubs that are not ranges or whose ranges overlap (then the generic hub kernel is used).
BG_DEVINL void mark_range_hub(int32_t b, const int32_t* __restrict__ colseg, int32_t n,
                              int32_t* __restrict__ hub_lo, int32_t* __restrict__ hub_of_row,
                              int32_t* __restrict__ info) {
  const int32_t lo = colseg[0];
  int ok = 1;
  for (int32_t i = threadIdx.x; i < n; i += blockDim.x) ok &= (colseg[i] == lo + i);
  ok = __syncthreads_and(ok);
  if (ok) {
    int clash = 0;
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) clash |= (atomicCAS(&hub_of_row[lo + i], -1, b) != -1);
    if (clash) atomicAdd(&info[4], 1);
  }
  if (threadIdx.x == 0) {
    hub_lo[b] = ok ? lo : -1;
    if (!ok) atomicAdd(&info[4], 1);
  }
}

// Sorts the n DISTINCT edge ids of a hub row with a bitmap: the ids of a hub's in-edges span a short interval of the edge
// list (the reference appends a graph's hub pairs (s, i), (i, s) in one block, VirtualEdgeCreate.py:106-111: span = 2n - 1),
// so "set bit id - min; rank = number of lower set bits" orders them in O(n + span / 32) instead of the bitonic
// network's O(n log^2 n).  `bits` holds span bits (span <= 32 * kSortSmemElems), `red` 64 words of scratch.
// Returns false (nothing written) when the span does not fit.
BG_DEVINL bool bitmap_sort_row(int32_t* __restrict__ seg, int32_t n, uint32_t* __restrict__ bits, int32_t* __restrict__ red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  int32_t lo = 0x7fffffff, hi = -1;
  for (int32_t i = threadIdx.x; i < n; i += blockDim.x) { const int32_t v = seg[i]; lo = min(lo, v); hi = max(hi, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if (lane == 0) { red[warp] = lo; red[32 + warp] = hi; }
  __syncthreads();
  lo = red[0]; hi = red[32];
  for (int w = 1; w < n_warps; ++w) { lo = min(lo, red[w]); hi = max(hi, red[32 + w]); }
  __syncthreads();
  const int64_t span = (int64_t)hi - lo + 1;
  const int32_t words = (int32_t)((span + 31) >> 5);
  if (span > (int64_t)32 * (kSortSmemElems - 64)) return false;                        // uniform over the CTA
  for (int32_t w = threadIdx.x; w < words; w += blockDim.x) bits[w] = 0u;
  __syncthreads();
  for (int32_t i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t d = (uint32_t)(seg[i] - lo); atomicOr(&bits[d >> 5], 1u << (d & 31)); }
  __syncthreads();
  // exclusive prefix of the per-word popcounts: each thread owns a contiguous run of words
  const int32_t per = (words + (int32_t)blockDim.x - 1) / (int32_t)blockDim.x;
  const int32_t w0 = min(words, (int32_t)threadIdx.x * per), w1 = min(words, w0 + per);
  int32_t mine = 0;
  for (int32_t w = w0; w < w1; ++w) mine += __popc(bits[w]);
  int32_t total;
  int32_t run = block_exclusive_scan(mine, red, total);
  __syncthreads();                                                               // every element of seg has been read
  for (int32_t w = w0; w < w1; ++w) {
    uint32_t b = bits[w];
    while (b) {
      const int bit = __ffs(b) - 1;
      b &= b - 1;
      seg[run++] = lo + (w << 5) + bit;
    }
  }
  __syncthreads();
  return true;
}

// one CTA per hub row (grid-stride over the list): bitmap sort of its edge ids (bitonic network as the fallback)
__global__ void __launch_bounds__(1024) k_csr_sort_big(const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ big_rows,
                                                       int32_t* __restrict__ info, int32_t max_big,
                                                       const int64_t* __restrict__ other, int64_t N,
                                                       int32_t* __restrict__ perm, int32_t* __restrict__ col,
                                                       int32_t* __restrict__ hub_lo, int32_t* __restrict__ hub_of_row) {
  extern __shared__ int32_t sbuf[];
  int32_t n_big = min(info[1], max_big);
  for (int32_t b = blockIdx.x; b < n_big; b += gridDim.x) {
    const int32_t r = big_rows[b];
    const int32_t s = rowptr[r], n = rowptr[r + 1] - s;
    int32_t P = 1;
    while (P < n) P <<= 1;
    if (bitmap_sort_row(perm + s, n, reinterpret_cast<uint32_t*>(sbuf) + 64, sbuf)) {
      for (int32_t i = threadIdx.x; i < n; i += blockDim.x) col[s + i] = checked_col(other[perm[s + i]], N, info);
      __syncthreads();
    } else if (n <= kSortSmemElems) {
      for (int32_t i = threadIdx.x; i < n; i += blockDim.x) sbuf[i] = perm[s + i];
      __syncthreads();
      bitonic_ascending(sbuf, n, P);
      for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
        int32_t p = sbuf[i];
        perm[s + i] = p;
        col[s + i] = checked_col(other[p], N, info);
      }
      __syncthreads();
    } else {
      bitonic_ascending(perm + s, n, P);   // global-memory fallback for giant rows
      for (int32_t i = threadIdx.x; i < n; i += blockDim.x) col[s + i] = checked_col(other[perm[s + i]], N, info);
      __syncthreads();
    }
    if (hub_of_row) mark_range_hub(b, col + s, n, hub_lo, hub_of_row, info);
    __syncthreads();
  }
}

// ------------------------------------------------------------------ batch vector -> graph offsets
__global__ void k_batch_info(const int64_t* __restrict__ batch, int64_t N, int32_t* __restrict__ info) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    int64_t b = batch[i];
    if (b < 0 || b >= 0x7fffffffLL) bad = true;
    if (i > 0 && batch[i - 1] > b) bad = true;
    if (i == N - 1) info[0] = (int32_t)(b + 1);
  }
  if (bad) atomicOr(&info[1], 1);
}

__global__ void k_graph_ptr(const int64_t* __restrict__ batch, int64_t N, int64_t G,
                            int32_t* __restrict__ graph_ptr) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += stride) {
    int64_t prev = (i == 0) ? -1 : batch[i - 1];
    int64_t cur = (i == N) ? G : batch[i];
    if (cur > G) cur = G;
    for (int64_t g = prev + 1; g <= cur; ++g) graph_ptr[g] = (int32_t)i;   // empty graphs get the same offset
  }
}

// ------------------------------------------------------------------ host entry points
struct CsrWorkspace {
  int32_t* deg;
  int32_t* cursor;
  int32_t* block_sums;
  int32_t n_scan_blocks;
  size_t bytes;
};

static inline CsrWorkspace csr_workspace_layout(void* base, int64_t N) {
  CsrWorkspace w;
  w.n_scan_blocks = (int32_t)ceil_div64(N > 0 ? N : 1, kScanItemsPerBlock);
  size_t off = 0;
  auto take = [&](size_t n_bytes) { size_t o = off; off += (n_bytes + 255) & ~(size_t)255; return o; };
  size_t o_deg = take(sizeof(int32_t) * (size_t)(N + 1));
  size_t o_cur = take(sizeof(int32_t) * (size_t)(N + 1));
  size_t o_bs = take(sizeof(int32_t) * (size_t)(w.n_scan_blocks + 1));
  w.bytes = off;
  char* b = static_cast<char*>(base);
  w.deg = reinterpret_cast<int32_t*>(b + o_deg);
  w.cursor = reinterpret_cast<int32_t*>(b + o_cur);
  w.block_sums = reinterpret_cast<int32_t*>(b + o_bs);
  return w;
}

}  // namespace bg
