// K3: tcgen05 tile GEMM with the SAGE update epilogue fused in.
//
//   out[m, 0:512] = epilogue( sum_s A_s[m, :] . B_s^T )       (up to 6 K-segments)
//
// The SAGE update is the two segments (agg, lin_l.weight) and (x, lin_r.weight), i.e.
// the concatenated [lin_l | lin_r] GEMM with K = 1024 -- replaces `lin_l(agg) +
// lin_r(x)`, `F.normalize`, eval `BatchNorm1d`, `ReLU` and the skip connection of the
// reference layer loop (Models/BuckGNN.py:449-457 + PyG SAGEConv.forward).  The same
// kernel runs the encoder's 128 -> 512 Linear (:73) as a one-segment GEMM.
//
// Structure (one persistent CTA per SM, or one CTA pair per SM pair with cta_group::2):
//   warp 0    TMA producer: A tile 128 x 128 B and the B (weight) rows of this CTA
//             into a ring of 128B-swizzled K-major stages
//   warp 1    allocates TMEM; one elected lane issues tcgen05.mma (M = 128*cg, N = 256,
//             two N halves -> all 512 output columns of a row tile live in TMEM)
//   warps 2-5 epilogue: one TMEM lane (= output row) per thread; pass 1 sums squares
//             of the whole 512-wide row (L2-normalize needs the full row, which is why
//             the tile spans all 512 columns = all 512 TMEM columns), pass 2 applies
//             bias / normalize / BN / ReLU / skip and stores.
// Tensor-core bound: 2*M*K*512 flops per tile row block; algorithmic bytes per row
// = (K_total + 512 [+512 residual]) * elem_size.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace bg {

constexpr int kTileM = 128;                       // rows per CTA
constexpr int kStageKBytes = 128;                 // one 128B swizzle row of K per stage
constexpr int kATileBytes = kTileM * kStageKBytes;        // 16 KB
constexpr int kGemmThreads = 192;
constexpr int kEpiParamBytes = 3 * kHidden * 4;   // bias, bn_scale, bn_shift staged in smem

template <int kCg> struct GemmCfg {
  static constexpr int kBRows = kHidden / kCg;                  // weight rows held by one CTA
  static constexpr int kBTileBytes = kBRows * kStageKBytes;     // 64 KB / 32 KB
  static constexpr int kStageBytes = kATileBytes + kBTileBytes; // 80 KB / 48 KB
  static constexpr int kStages = (kCg == 1) ? 2 : 4;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiParamBytes + 256;
};

struct alignas(64) GemmSeg { CUtensorMap a; CUtensorMap b; };

struct alignas(64) GemmParams {
  GemmSeg seg[BG_MAX_GEMM_SEGMENTS];
  int32_t kblocks[BG_MAX_GEMM_SEGMENTS];
  int32_t n_seg;
  int32_t k_elems_per_block;        // 64 (16-bit operands) or 32 (tf32)
  uint32_t a_fmt, b_fmt;            // UMMA operand formats: 0 f16, 1 bf16, 2 tf32
  int32_t n_tiles;                  // row tiles of 128*cg rows
  int32_t normalize, relu;
  int64_t m;
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  const void* residual;
  int64_t ldr;
  void* out;
  int64_t ldo;
};

enum : uint32_t { kTagEmpty = 1, kTagFull = 2, kTagTmemEmpty = 3, kTagTmemFull = 4 };

template <int kCg, bool kTf32, typename TOut>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm512(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<kCg>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t gemm_smem_raw[];
  const uint32_t smem_base = (smem_u32(gemm_smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024 B
  uint8_t* smem_gen = gemm_smem_raw + (smem_base - smem_u32(gemm_smem_raw));
  const uint32_t stages_u32 = smem_base;
  float* epi_params = reinterpret_cast<float*>(smem_gen + kStages * Cfg::kStageBytes);
  const uint32_t bars_u32 = stages_u32 + kStages * Cfg::kStageBytes + kEpiParamBytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + kStages * Cfg::kStageBytes + kEpiParamBytes + 192);
  auto full_bar = [&](int s) { return bars_u32 + 8u * s; };
  auto empty_bar = [&](int s) { return bars_u32 + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bars_u32 + 8u * (2 * kStages);
  const uint32_t tmem_empty_bar = bars_u32 + 8u * (2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCg == 2) ? cluster_ctarank() : 0u;
  const int tile0 = (kCg == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int tile_stride = (kCg == 2) ? (gridDim.x >> 1) : gridDim.x;

  // ---- one-time setup
  for (int i = threadIdx.x; i < kHidden; i += kGemmThreads) {
    epi_params[i] = p.bias ? p.bias[i] : 0.f;
    epi_params[kHidden + i] = p.bn_scale ? p.bn_scale[i] : 1.f;
    epi_params[2 * kHidden + i] = p.bn_scale ? p.bn_shift[i] : 0.f;
  }
  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.n_seg; ++s) { tma_prefetch_desc(&p.seg[s].a); tma_prefetch_desc(&p.seg[s].b); }
  }
  if (kCg == 2) cluster_sync();                // both CTAs resident before the paired TMEM alloc
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), kCg); mbar_init(empty_bar(s), 1); }
      mbar_init(tmem_full_bar, 1);
      mbar_init(tmem_empty_bar, kCg * 128);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kCg>(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  if (kCg == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int tile = tile0; tile < p.n_tiles; tile += tile_stride) {
        const int32_t row0 = tile * (kTileM * kCg) + (int32_t)rank * kTileM;
        for (int s = 0; s < p.n_seg; ++s) {
          const void* map_a = &p.seg[s].a;
          const void* map_b = &p.seg[s].b;
          for (int kb = 0; kb < p.kblocks[s]; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, kTagEmpty);
            const uint32_t sa = stages_u32 + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + kATileBytes;
            const int32_t k0 = kb * p.k_elems_per_block;
            if constexpr (kCg == 1) {
              mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_2d(sa, map_a, full_bar(stage), k0, row0);
#pragma unroll
              for (int j = 0; j < Cfg::kBRows / 128; ++j)
                tma_load_2d(sb + j * 16384, map_b, full_bar(stage), k0, j * 128);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
              else mbar_arrive_cluster(full_bar(stage), 0);
              tma_load_2d_cg2(sa, map_a, full_bar(stage), k0, row0);
#pragma unroll
              for (int j = 0; j < Cfg::kBRows / 128; ++j)      // N half j: weight rows j*256 + rank*128
                tma_load_2d_cg2(sb + j * 16384, map_b, full_bar(stage), k0, j * 256 + (int32_t)rank * 128);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc(p.a_fmt, p.b_fmt, kTileM * kCg, 256);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = tile0; tile < p.n_tiles; tile += tile_stride, ++it) {
        mbar_wait(tmem_empty_bar, (it & 1u) ^ 1u, kTagTmemEmpty);   // epilogue drained the accumulator
        tc_fence_after();
        uint32_t first = 1;
        for (int s = 0; s < p.n_seg; ++s) {
          const int nkb = p.kblocks[s];
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full_bar(stage), phase, kTagFull);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = stages_u32 + stage * Cfg::kStageBytes;
              const uint32_t sb = sa + kATileBytes;
#pragma unroll
              for (int k = 0; k < kStageKBytes / 32; ++k) {
                const uint64_t da = umma_smem_desc(sa + k * 32);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint64_t db = umma_smem_desc(sb + h * (Cfg::kBTileBytes / 2) + k * 32);
                  umma<kCg, kTf32>(tmem_base + h * 256, da, db, idesc, (first && k == 0) ? 0u : 1u);
                }
              }
              umma_commit<kCg>(empty_bar(stage));                   // smem slot reusable once these MMAs retire
              if (s == p.n_seg - 1 && kb == nkb - 1) umma_commit<kCg>(tmem_full_bar);
            }
            __syncwarp();
            first = 0;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================ epilogue warps 2..5
    const int q = warp & 3;                                     // TMEM lane quarter this warp may read
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* sbias = epi_params;
    const float* sscale = epi_params + kHidden;
    const float* sshift = epi_params + 2 * kHidden;
    const bool has_bn = p.bn_scale != nullptr;
    uint32_t it = 0;
    for (int tile = tile0; tile < p.n_tiles; tile += tile_stride, ++it) {
      const int64_t m = (int64_t)tile * (kTileM * kCg) + rank * kTileM + q * 32 + lane;
      const bool valid = m < p.m;
      mbar_wait(tmem_full_bar, it & 1u, kTagTmemFull);
      tc_fence_after();
      float inv = 1.f;
      if (p.normalize) {
        float ss = 0.f;
        for (int c0 = 0; c0 < kHidden; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { float v = __uint_as_float(r[i]) + sbias[c0 + i]; ss = fmaf(v, v, ss); }
        }
        inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      }
      TOut* orow = reinterpret_cast<TOut*>(p.out) + m * p.ldo;
      const TOut* rrow = p.residual ? reinterpret_cast<const TOut*>(p.residual) + m * p.ldr : nullptr;
      for (int c0 = 0; c0 < kHidden; c0 += 32) {
        constexpr int kVec = 32 * sizeof(TOut) / 16;             // 16-byte vectors per 32 columns
        uint4 res[kVec];
        if (rrow && valid) {
#pragma unroll
          for (int j = 0; j < kVec; ++j) res[j] = ldg_nc_v4(reinterpret_cast<const uint4*>(rrow + c0) + j);
        }
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = (__uint_as_float(r[i]) + sbias[c0 + i]) * inv;
          if (has_bn) t = fmaf(t, sscale[c0 + i], sshift[c0 + i]);
          if (p.relu) t = fmaxf(t, 0.f);
          v[i] = t;
        }
        if (rrow && valid) {
          const uint32_t* ru = reinterpret_cast<const uint32_t*>(res);
          if constexpr (sizeof(TOut) == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[2 * i] += Pack16<TOut>::lo(ru[i]); v[2 * i + 1] += Pack16<TOut>::hi(ru[i]); }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(ru[i]);
          }
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(orow + c0);
          if constexpr (sizeof(TOut) == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = Pack16<TOut>::pack(v[8 * j], v[8 * j + 1]); o.y = Pack16<TOut>::pack(v[8 * j + 2], v[8 * j + 3]);
              o.z = Pack16<TOut>::pack(v[8 * j + 4], v[8 * j + 5]); o.w = Pack16<TOut>::pack(v[8 * j + 6], v[8 * j + 7]);
              stg_v4(dst + j, o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 o;
              o.x = __float_as_uint(v[4 * j]); o.y = __float_as_uint(v[4 * j + 1]);
              o.z = __float_as_uint(v[4 * j + 2]); o.w = __float_as_uint(v[4 * j + 3]);
              stg_v4(dst + j, o);
            }
          }
        }
      }
      // accumulator fully read: hand TMEM back to the MMA warp of the leader CTA
      tc_fence_before();
      if (kCg == 1 || rank == 0) mbar_arrive(tmem_empty_bar);
      else mbar_arrive_cluster(tmem_empty_bar, 0);
    }
  }

  // ---- teardown
  tc_fence_before();
  if (kCg == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) tmem_dealloc<kCg>(tmem_base, 512);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tensorMapEncodeTiled get_tensor_map_encoder();

// [rows, k] row-major matrix, box = 128 rows x 128 bytes of K, 128B swizzle, zero fill out of bounds
// fmt: UMMA operand format (0 f16, 1 bf16, 2 tf32/f32)
static inline int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int64_t ld,
                                   uint32_t fmt) {
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return BG_ERR_CUDA;
  const bool tf32 = fmt == 2;
  const int esz = tf32 ? 4 : 2;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(kStageKBytes / esz), 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, dt, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BG_OK : BG_ERR_CUDA;
}

template <int kCg, bool kTf32, typename TOut>
static int launch_gemm512(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<kCg>;
  auto kern = k_gemm512<kCg, kTf32, TOut>;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int sms = sm_count();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (kCg == 1) {
    cfg.gridDim = dim3((unsigned)min(p.n_tiles, sms));
    cfg.numAttrs = 0;
  } else {
    cfg.gridDim = dim3((unsigned)(2 * min(p.n_tiles, sms / 2)));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
  }
  cfg.attrs = attr;
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  BG_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  return BG_OK;
}

}  // namespace bg
