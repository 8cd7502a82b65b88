// K5 (front): first two Linear+ReLU of `node_encoder`, fused (reference
// Models/BuckGNN.py:68-72, applied at :323):  h = relu(relu(x W1^T + b1) W2^T + b2),
// x [N,F<=32] f32 -> h [N,128].  The third Linear (128 -> 512) runs on the tensor-core
// GEMM (gemm_tc.cuh) with h as its A operand.
//
// K = F (16) and K = 64 are too thin for tcgen05 tiles to pay, so this is a register-
// tiled CUDA-core kernel: 64-row tiles, weights resident in shared memory for the
// CTA's lifetime, fp32 FMAs (exactly the reference arithmetic, no operand rounding).
#pragma once
#include "common.cuh"

namespace bg {

constexpr int kEncRows = 64;       // rows per tile
constexpr int kEncThreads = 256;
constexpr int kEncH1 = 64;
constexpr int kEncH2 = 128;
constexpr int kEncMaxF = 32;

struct EncoderSmem {
  float w1t[kEncMaxF][kEncH1];         // [k][n]
  float b1[kEncH1];
  float w2t[kEncH1][kEncH2];           // [k][n]
  float b2[kEncH2];
  float xt[kEncMaxF][kEncRows];        // [k][m]
  float h1t[kEncH1][kEncRows + 4];     // [k][m]
};

template <typename TOut>
__global__ void __launch_bounds__(kEncThreads)
k_encoder_front(const float* __restrict__ x, int64_t N, int F,
                const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2,
                const int32_t* __restrict__ row_gather, TOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char enc_smem_raw[];
  EncoderSmem& s = *reinterpret_cast<EncoderSmem*>(enc_smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < kEncH1 * F; i += kEncThreads) { int n = i / F, k = i % F; s.w1t[k][n] = w1[i]; }
  for (int i = tid; i < kEncH2 * kEncH1; i += kEncThreads) { int n = i / kEncH1, k = i % kEncH1; s.w2t[k][n] = w2[i]; }
  if (tid < kEncH1) s.b1[tid] = b1[tid];
  if (tid < kEncH2) s.b2[tid] = b2[tid];
  __syncthreads();

  const int64_t n_tiles = (N + kEncRows - 1) / kEncRows;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 thread grid
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t m0 = tile * kEncRows;
    // stage x tile transposed: xt[k][m]
    for (int i = tid; i < kEncRows * F; i += kEncThreads) {
      int m = i / F, k = i % F;
      float val = 0.f;
      if (m0 + m < N) {
        const int64_t src = row_gather ? (int64_t)row_gather[m0 + m] : (m0 + m);
        val = x[src * F + k];
      }
      s.xt[k][m] = val;
    }
    __syncthreads();
    {  // layer 1: thread -> rows ty*4..+4, cols tx*4..+4
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = s.b1[tx * 4 + b];
      for (int k = 0; k < F; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(&s.xt[k][ty * 4]);
        const float4 wv = *reinterpret_cast<const float4*>(&s.w1t[k][tx * 4]);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], wa[b], acc[a][b]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a) s.h1t[tx * 4 + b][ty * 4 + a] = fmaxf(acc[a][b], 0.f);
    }
    __syncthreads();
    {  // layer 2: thread -> rows ty*4..+4, cols tx*8..+8
      float acc[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = s.b2[tx * 8 + b];
#pragma unroll 4
      for (int k = 0; k < kEncH1; ++k) {
        const float4 hv = *reinterpret_cast<const float4*>(&s.h1t[k][ty * 4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&s.w2t[k][tx * 8]);
        const float4 w1v = *reinterpret_cast<const float4*>(&s.w2t[k][tx * 8 + 4]);
        const float ha[4] = {hv.x, hv.y, hv.z, hv.w};
        const float wa[8] = {w0.x, w0.y, w0.z, w0.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(ha[a], wa[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int64_t m = m0 + ty * 4 + a;
        if (m >= N) continue;
        float v[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) v[b] = fmaxf(acc[a][b], 0.f);
        if constexpr (sizeof(TOut) == 2) {
          uint4 q;
          q.x = Pack16<TOut>::pack(v[0], v[1]); q.y = Pack16<TOut>::pack(v[2], v[3]);
          q.z = Pack16<TOut>::pack(v[4], v[5]); q.w = Pack16<TOut>::pack(v[6], v[7]);
          stg_v4(out + m * kEncH2 + tx * 8, q);
        } else {
          uint4 q0, q1;
          q0.x = __float_as_uint(v[0]); q0.y = __float_as_uint(v[1]); q0.z = __float_as_uint(v[2]); q0.w = __float_as_uint(v[3]);
          q1.x = __float_as_uint(v[4]); q1.y = __float_as_uint(v[5]); q1.z = __float_as_uint(v[6]); q1.w = __float_as_uint(v[7]);
          stg_v4(out + m * kEncH2 + tx * 8, q0);
          stg_v4(out + m * kEncH2 + tx * 8 + 4, q1);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace bg
