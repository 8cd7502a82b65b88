// K4: segmented global mean pool + regression head.
//
// Replaces `global_mean_pool(x, batch)` and `decoder(pooled).squeeze()` of the
// reference (Models/BuckGNN.py:274, 515-516; decoder shapes :94-100).  HBM bound:
// one read pass over x [N,512].
//   k_pool_partial : grid (G, kPoolSlices); each CTA sums a contiguous slice of one
//                    graph's rows (warp per row subset, fp32), writes partial[g][s][512]
//   k_pool_head    : one CTA per graph: adds the slices in order, divides by
//                    max(count,1), runs Linear(512,128) ReLU Linear(128,64) ReLU
//                    Linear(64,out_dim) in fp32.
// Deterministic: no atomics, fixed reduction order.
#pragma once
#include "common.cuh"
#include "aggregate.cuh"

namespace bg {

constexpr int kPoolSlices = 8;
constexpr int kPoolWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kPoolWarps * 32)
k_pool_partial(const T* __restrict__ x, const int32_t* __restrict__ graph_ptr, float* __restrict__ partial) {
  __shared__ float red[kPoolWarps][kHidden];
  const int g = blockIdx.x, slice = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t beg = graph_ptr[g], cnt = graph_ptr[g + 1] - beg;
  const int32_t s_beg = beg + (int32_t)((int64_t)cnt * slice / kPoolSlices);
  const int32_t s_end = beg + (int32_t)((int64_t)cnt * (slice + 1) / kPoolSlices);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  int32_t r = s_beg + warp;
  for (; r + kPoolWarps < s_end; r += 2 * kPoolWarps) {      // two rows in flight per warp
    RowFrag<T> f0, f1;
    f0.load(x + (size_t)r * kHidden, lane);
    f1.load(x + (size_t)(r + kPoolWarps) * kHidden, lane);
    f0.template accumulate<BG_AGGR_SUM>(acc);
    f1.template accumulate<BG_AGGR_SUM>(acc);
  }
  for (; r < s_end; r += kPoolWarps) {
    RowFrag<T> f;
    f.load(x + (size_t)r * kHidden, lane);
    f.template accumulate<BG_AGGR_SUM>(acc);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) red[warp][RowFrag<T>::col_of(lane, i)] = acc[i];
  __syncthreads();
  float* out = partial + ((size_t)g * kPoolSlices + slice) * kHidden;
  for (int c = threadIdx.x; c < kHidden; c += blockDim.x) {
    float v = red[0][c];
#pragma unroll
    for (int w = 1; w < kPoolWarps; ++w) v += red[w][c];
    out[c] = v;
  }
}

__global__ void __launch_bounds__(128)
k_pool_head(const float* __restrict__ partial, const int32_t* __restrict__ graph_ptr,
            const float* __restrict__ w1, const float* __restrict__ b1,
            const float* __restrict__ w2, const float* __restrict__ b2,
            const float* __restrict__ w3, const float* __restrict__ b3, int out_dim,
            float* __restrict__ pred, float* __restrict__ pooled_out) {
  __shared__ float pooled[kHidden];
  __shared__ float h1[128];
  __shared__ float h2[64];
  const int g = blockIdx.x, t = threadIdx.x;
  const float cnt = (float)max(graph_ptr[g + 1] - graph_ptr[g], 1);
  for (int c = t; c < kHidden; c += 128) {
    const float* pp = partial + (size_t)g * kPoolSlices * kHidden + c;
    float v = pp[0];
#pragma unroll
    for (int s = 1; s < kPoolSlices; ++s) v += pp[(size_t)s * kHidden];
    v = v / cnt;
    pooled[c] = v;
    if (pooled_out) pooled_out[(size_t)g * kHidden + c] = v;
  }
  __syncthreads();
  {  // 512 -> 128: one output per thread, 4 independent partial sums
    const float4* wr = reinterpret_cast<const float4*>(w1 + (size_t)t * kHidden);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = 0; k < kHidden / 4; ++k) {
      const float4 w = __ldg(wr + k);
      a0 = fmaf(w.x, pooled[4 * k], a0); a1 = fmaf(w.y, pooled[4 * k + 1], a1);
      a2 = fmaf(w.z, pooled[4 * k + 2], a2); a3 = fmaf(w.w, pooled[4 * k + 3], a3);
    }
    h1[t] = fmaxf((a0 + a1) + (a2 + a3) + b1[t], 0.f);
  }
  __syncthreads();
  if (t < 64) {
    const float* wr = w2 + (size_t)t * 128;
    float a = 0.f;
    for (int k = 0; k < 128; ++k) a = fmaf(__ldg(wr + k), h1[k], a);
    h2[t] = fmaxf(a + b2[t], 0.f);
  }
  __syncthreads();
  if (t < out_dim) {
    const float* wr = w3 + (size_t)t * 64;
    float a = 0.f;
    for (int k = 0; k < 64; ++k) a = fmaf(__ldg(wr + k), h2[k], a);
    pred[(size_t)g * out_dim + t] = a + b3[t];
  }
}

}  // namespace bg
