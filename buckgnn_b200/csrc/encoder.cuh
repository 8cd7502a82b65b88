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


// ------------------------------------------------------------------ tensor-core variant (16-bit outputs)
// Same two layers on mma.sync m16n8k16 (f16 operands, fp32 accumulate): K = 16 and K = 64 are far too thin
// for a tcgen05/TMEM pipeline, but the warp-level MMA turns the LDS-bound FMA loop above (0.73 ms at cfg 2)
// into a kernel bound by its 0.33 GB of HBM traffic.  One warp owns 16 rows: layer 1's accumulator fragments
// ARE layer 2's A fragments (C tiles 2k, 2k+1 -> A k-block k), so h1 never leaves registers; the weights sit
// in shared memory as fp16 with padded rows (conflict-free fragment loads); the output tile is staged through
// shared memory for 256-byte coalesced row stores.  Operand rounding: x, W1, h1, W2 to fp16 (11-bit
// significands, like the fp16 GEMM operands downstream); the fp32 / tf32 precision modes keep the exact kernel.
constexpr int kEncP1 = 40;         // halves per W1 row in smem (K padded to 32, +8 against bank conflicts)
constexpr int kEncP2 = 72;         // halves per W2 row (64 + 8)
constexpr int kEncPO = 136;        // halves per staged output row (128 + 8)
constexpr int kEncMmaWarps = 8;

struct EncoderMmaSmem {
  __half w1[kEncH1][kEncP1];
  __half w2[kEncH2][kEncP2];
  float b1[kEncH1];
  float b2[kEncH2];
  __half stage[kEncMmaWarps][16][kEncPO];   // reinterpreted as TOut (same size)
};

BG_DEVINL void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename TOut>
__global__ void __launch_bounds__(kEncMmaWarps * 32)
k_encoder_front_mma(const float* __restrict__ x, int64_t N, int F,
                    const float* __restrict__ w1, const float* __restrict__ b1,
                    const float* __restrict__ w2, const float* __restrict__ b2,
                    const int32_t* __restrict__ row_gather, TOut* __restrict__ out, int32_t* __restrict__ nonfinite) {
  static_assert(sizeof(TOut) == 2, "16-bit outputs only");
  // h is stored BEFORE any normalisation: a checkpoint with large encoder activations would overflow the fp16 range
  // silently (inf -> NaN in the L2-normalize -> 0 after ReLU's fmaxf).  Values the output format cannot hold raise
  // *nonfinite instead; bg_pool_head turns the flag into NaN predictions (no host sync involved).
  const float limit = is_bf16<TOut>::value ? 3.3e38f : 65504.f;
  bool bad = false;
  extern __shared__ __align__(16) unsigned char enc_smem_raw[];
  EncoderMmaSmem& s = *reinterpret_cast<EncoderMmaSmem*>(enc_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kEncH1 * kEncP1; i += blockDim.x) {
    const int n = i / kEncP1, k = i % kEncP1;
    s.w1[n][k] = __float2half_rn(k < F ? w1[n * F + k] : 0.f);
  }
  for (int i = tid; i < kEncH2 * kEncH1; i += blockDim.x) s.w2[i / kEncH1][i % kEncH1] = __float2half_rn(w2[i]);
  if (tid < kEncH1) s.b1[tid] = b1[tid];
  if (tid < kEncH2) s.b2[tid] = b2[tid];
  __syncthreads();

  const int g = lane >> 2, tig = lane & 3;
  const int ksteps1 = (F + 15) / 16;
  const int64_t n_tiles = (N + 15) / 16;
  TOut (*stage)[kEncPO] = reinterpret_cast<TOut (*)[kEncPO]>(s.stage[warp]);
  for (int64_t tile = (int64_t)blockIdx.x * kEncMmaWarps + warp; tile < n_tiles; tile += (int64_t)gridDim.x * kEncMmaWarps) {
    const int64_t m0 = tile * 16;
    const int64_t r0 = m0 + g, r1 = m0 + g + 8;
    const float* x0 = nullptr;
    const float* x1 = nullptr;
    if (r0 < N) x0 = x + (size_t)(row_gather ? (int64_t)row_gather[r0] : r0) * F;
    if (r1 < N) x1 = x + (size_t)(row_gather ? (int64_t)row_gather[r1] : r1) * F;
    auto ld = [&](const float* row, int k) { return (row != nullptr && k < F) ? row[k] : 0.f; };
    // ---- layer 1: C1[j] = x W1^T + b1, 8 n-tiles of 8 columns
    float c1[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ba = s.b1[8 * j + 2 * tig], bb = s.b1[8 * j + 2 * tig + 1];
      c1[j][0] = ba; c1[j][1] = bb; c1[j][2] = ba; c1[j][3] = bb;
    }
    for (int ks = 0; ks < ksteps1; ++ks) {
      const int k0 = 16 * ks + 2 * tig;
      uint32_t a[4];
      a[0] = pack_f16(ld(x0, k0), ld(x0, k0 + 1));
      a[1] = pack_f16(ld(x1, k0), ld(x1, k0 + 1));
      a[2] = pack_f16(ld(x0, k0 + 8), ld(x0, k0 + 9));
      a[3] = pack_f16(ld(x1, k0 + 8), ld(x1, k0 + 9));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t b0v = *reinterpret_cast<const uint32_t*>(&s.w1[8 * j + g][k0]);
        const uint32_t b1v = *reinterpret_cast<const uint32_t*>(&s.w1[8 * j + g][k0 + 8]);
        mma_16816(c1[j], a, b0v, b1v);
      }
    }
    // ---- ReLU, and the accumulator fragments become layer 2's A fragments (4 k-blocks of 16)
    uint32_t a2[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      a2[kk][0] = pack_f16(fmaxf(c1[2 * kk][0], 0.f), fmaxf(c1[2 * kk][1], 0.f));
      a2[kk][1] = pack_f16(fmaxf(c1[2 * kk][2], 0.f), fmaxf(c1[2 * kk][3], 0.f));
      a2[kk][2] = pack_f16(fmaxf(c1[2 * kk + 1][0], 0.f), fmaxf(c1[2 * kk + 1][1], 0.f));
      a2[kk][3] = pack_f16(fmaxf(c1[2 * kk + 1][2], 0.f), fmaxf(c1[2 * kk + 1][3], 0.f));
    }
    // ---- layer 2: 16 n-tiles, two at a time
#pragma unroll 2
    for (int j = 0; j < 16; j += 2) {
      float ca[4], cb[4];
      {
        const float ba = s.b2[8 * j + 2 * tig], bb = s.b2[8 * j + 2 * tig + 1];
        ca[0] = ba; ca[1] = bb; ca[2] = ba; ca[3] = bb;
        const float bc = s.b2[8 * j + 8 + 2 * tig], bd = s.b2[8 * j + 8 + 2 * tig + 1];
        cb[0] = bc; cb[1] = bd; cb[2] = bc; cb[3] = bd;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k0 = 16 * kk + 2 * tig;
        mma_16816(ca, a2[kk], *reinterpret_cast<const uint32_t*>(&s.w2[8 * j + g][k0]),
                  *reinterpret_cast<const uint32_t*>(&s.w2[8 * j + g][k0 + 8]));
        mma_16816(cb, a2[kk], *reinterpret_cast<const uint32_t*>(&s.w2[8 * j + 8 + g][k0]),
                  *reinterpret_cast<const uint32_t*>(&s.w2[8 * j + 8 + g][k0 + 8]));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) bad |= !(ca[e] <= limit) | !(cb[e] <= limit);          // NaN compares false
      *reinterpret_cast<uint32_t*>(&stage[g][8 * j + 2 * tig]) = Pack16<TOut>::pack(fmaxf(ca[0], 0.f), fmaxf(ca[1], 0.f));
      *reinterpret_cast<uint32_t*>(&stage[g + 8][8 * j + 2 * tig]) = Pack16<TOut>::pack(fmaxf(ca[2], 0.f), fmaxf(ca[3], 0.f));
      *reinterpret_cast<uint32_t*>(&stage[g][8 * j + 8 + 2 * tig]) = Pack16<TOut>::pack(fmaxf(cb[0], 0.f), fmaxf(cb[1], 0.f));
      *reinterpret_cast<uint32_t*>(&stage[g + 8][8 * j + 8 + 2 * tig]) = Pack16<TOut>::pack(fmaxf(cb[2], 0.f), fmaxf(cb[3], 0.f));
    }
    __syncwarp();
    // ---- coalesced store: 2 rows x 256 B per instruction
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = 2 * it + (lane >> 4), ch = lane & 15;
      if (m0 + r < N) stg_v4(out + (size_t)(m0 + r) * kEncH2 + ch * 8, *reinterpret_cast<const uint4*>(&stage[r][ch * 8]));
    }
    __syncwarp();
  }
  if (bad && nonfinite != nullptr) atomicOr(nonfinite, 1);
}

}  // namespace bg
