// Backward of the 'max' neighbourhood aggregation (SAGEConv(aggr='max'), Models/BuckGNN.py:171-176, 459-471) for the
// training step.
//
// Forward: agg_i[c] = max_{j -> i} x_j[c]  (0 for a node without in-edges).  The gradient follows what autograd gives
// the reference when PyG reduces with torch's `scatter_reduce_(..., 'amax', include_self=False)` on a zero-initialised
// output (its path without torch_scatter): the upstream gradient of (i, c) is shared EVENLY by all neighbours that
// attain the maximum, and the zero-initialised output element counts as one more sharer when the maximum is 0 --
//     n_i[c]  = [agg_i[c] == 0] + #{j -> i : x_j[c] == agg_i[c]}
//     dx_j[c] = sum_{i : j -> i} [x_j[c] == agg_i[c]] * dagg_i[c] / n_i[c]
// (after a ReLU ties at 0 are the common case, so this detail matters).  Two gather passes, no atomics:
//   kMode 0 over the CSR keyed by TARGET:  w_i = dagg_i / n_i                       (ref = agg_i, neighbours x_j)
//   kMode 1 over the CSR keyed by SOURCE:  dx_j = sum_i [agg_i == x_j] w_i          (ref = x_j,  neighbours agg_i, w_i)
// Rows up to the big-row threshold: one warp per row, two neighbour rows in flight.  Hub rows: one CTA per row, its
// warps take interleaved neighbours and their partial sums are added in a fixed order.
#pragma once
#include "common.cuh"
#include "aggregate.cuh"
#include "train.cuh"

namespace bg {

constexpr int kMaxBwdWarps = 8;

// acc += [b == ref] * (kMode == 0 ? 1 : w) over neighbours col[beg + first], col[beg + first + step], ...
template <typename T, int kMode>
BG_DEVINL void match_accumulate(const float (&ref)[16], const T* __restrict__ b, const T* __restrict__ w,
                                const int32_t* __restrict__ col, int32_t beg, int32_t end, int first, int step, int lane,
                                float (&acc)[16]) {
  int32_t e = beg + first;
  for (; e + step < end; e += 2 * step) {                     // two neighbour rows in flight
    const int32_t k0 = col[e], k1 = col[e + step];
    float v0[16], v1[16];
    row_load<T>(b + (size_t)k0 * kHidden, lane, v0);
    row_load<T>(b + (size_t)k1 * kHidden, lane, v1);
    if constexpr (kMode == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] += (v0[i] == ref[i] ? 1.f : 0.f) + (v1[i] == ref[i] ? 1.f : 0.f);
    } else {
      float w0[16], w1[16];
      row_load<T>(w + (size_t)k0 * kHidden, lane, w0);
      row_load<T>(w + (size_t)k1 * kHidden, lane, w1);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] += (v0[i] == ref[i] ? w0[i] : 0.f) + (v1[i] == ref[i] ? w1[i] : 0.f);
    }
  }
  for (; e < end; e += step) {
    const int32_t k0 = col[e];
    float v0[16];
    row_load<T>(b + (size_t)k0 * kHidden, lane, v0);
    if constexpr (kMode == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] += (v0[i] == ref[i] ? 1.f : 0.f);
    } else {
      float w0[16];
      row_load<T>(w + (size_t)k0 * kHidden, lane, w0);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] += (v0[i] == ref[i] ? w0[i] : 0.f);
    }
  }
}

// kMode 0: out_r = d_r / (acc + [ref == 0]);   kMode 1: out_r = acc
template <typename T, int kMode>
BG_DEVINL void match_finish(const float (&ref)[16], const T* __restrict__ d_row, float (&acc)[16], int lane) {
  if constexpr (kMode == 0) {
    float d[16];
    row_load<T>(d_row, lane, d);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = d[i] / (acc[i] + (ref[i] == 0.f ? 1.f : 0.f));
  }
}

template <typename T, int kMode>
__global__ void __launch_bounds__(kMaxBwdWarps * 32)
k_max_bwd_rows(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ w, const T* __restrict__ d,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t N, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < N; r += n_warps) {
    const int32_t beg = rowptr[r], end = rowptr[r + 1];
    if (end - beg > kBigRowThreshold) continue;                // k_max_bwd_big
    float ref[16], acc[16];
    row_load<T>(a + (size_t)r * kHidden, lane, ref);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    match_accumulate<T, kMode>(ref, b, w, col, beg, end, 0, 1, lane, acc);
    match_finish<T, kMode>(ref, d + (size_t)r * kHidden, acc, lane);
    RowFrag<T>::store(out + (size_t)r * kHidden, lane, acc);
  }
}

template <typename T, int kMode>
__global__ void __launch_bounds__(kMaxBwdWarps * 32)
k_max_bwd_big(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ w, const T* __restrict__ d,
              const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ big_rows,
              T* __restrict__ out) {
  __shared__ float red[kMaxBwdWarps][kHidden];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t r = big_rows[blockIdx.x];
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  float ref[16], acc[16];
  row_load<T>(a + (size_t)r * kHidden, lane, ref);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  match_accumulate<T, kMode>(ref, b, w, col, beg, end, warp, kMaxBwdWarps, lane, acc);
#pragma unroll
  for (int i = 0; i < 16; ++i) red[warp][RowFrag<T>::col_of(lane, i)] = acc[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = RowFrag<T>::col_of(lane, i);
      float v = red[0][c];
#pragma unroll
      for (int k = 1; k < kMaxBwdWarps; ++k) v += red[k][c];
      acc[i] = v;
    }
    match_finish<T, kMode>(ref, d + (size_t)r * kHidden, acc, lane);
    RowFrag<T>::store(out + (size_t)r * kHidden, lane, acc);
  }
}

}  // namespace bg
